"""Accuracy of the panel-moment evaluation of the spike-time sums (include/svgpfa_b200.h, SVGPFA_SPIKE_PANEL), in numpy:
   sum_s c_s f(t_s)  against  sum_i m_i f(t_i),  m_i = sum_s c_s l_i(t_s)   (16 first-kind Chebyshev nodes per panel)
for f = kappa(. - z_j), dkappa/ddelta, dkappa/dlengthscale of both kernels, 10^4 spikes with random weights on [0, 1].
The second table sweeps the panel count: it calibrates the rule `panel half-width <= beta x scale of variation` that
B200SVLowerBound._panels_required applies (beta = 0.625 exponential-quadratic, 0.55 periodic with scale
p min(l, 1) / (2 pi)).      python tools/panel_accuracy.py"""
import numpy as np
rng = np.random.default_rng(0)
def lagrange_weights(x, p):
    # x in [-1,1], returns (len(x), p) weights at first-kind Chebyshev nodes
    i = np.arange(p); xi = np.cos(np.pi*(i+0.5)/p)
    n = np.arange(p)
    Tn_x = np.cos(np.outer(np.arccos(np.clip(x,-1,1)), n))        # (S,p)
    Tn_xi = np.cos(np.outer(np.arccos(xi), n))                    # (p,p)  [i,n]
    w = np.ones(p); w[0] = 0.5
    return (2.0/p) * (Tn_x * w) @ Tn_xi.T, xi
def nodes(lo, hi, B, p):
    i = np.arange(p); xi = np.cos(np.pi*(i+0.5)/p)
    wdt = (hi-lo)/B
    c = lo + (np.arange(B)+0.5)*wdt
    return (c[:,None] + 0.5*wdt*xi[None,:]).reshape(-1)
def moments(t, c, lo, hi, B, p):
    wdt = (hi-lo)/B
    b = np.clip(np.floor((t-lo)/wdt).astype(int), 0, B-1)
    x = (t - (lo + (b+0.5)*wdt)) / (0.5*wdt)
    L, _ = lagrange_weights(x, p)
    m = np.zeros((B,p))
    np.add.at(m, b, L*c[:,None])
    return m.reshape(-1)
def kern(kind, d, l, P=None):
    if kind == "eq":
        k = np.exp(-0.5*d**2/l**2); return k, k*(-d/l**2), k*d**2/l**3
    s = np.sin(np.pi*d/P); k = np.exp(-2*s**2/l**2)
    return k, k*(-2*np.pi/(P*l**2))*np.sin(2*np.pi*d/P), k*4*s**2/l**3
S = 10000; T = 1.0
t = np.sort(rng.uniform(0, T, S)); c = 0.3*rng.standard_normal(S)
z = np.linspace(0, T, 32) + rng.uniform(-0.1,0.1,32)*T/32
for kind, l, P in [("eq",0.1,None),("eq",0.15,None),("eq",0.3,None),("eq",1.05,None),("eq",0.05,None),("per",1.1,0.55),("per",1.9,0.95),("per",0.4,1.5)]:
    ex = [ (c[:,None]*f).sum(0) for f in kern(kind, t[:,None]-z[None,:], l, P)]
    scale = [np.abs(c).sum()*np.abs(f).max() for f in kern(kind, t[:,None]-z[None,:], l, P)]
    for B,p in [(4,16),(8,16),(8,20),(16,16),(16,12),(32,12)]:
        tn = nodes(0,T,B,p); m = moments(t,c,0,T,B,p)
        ap = [ (m[:,None]*f).sum(0) for f in kern(kind, tn[:,None]-z[None,:], l, P)]
        errs = [np.abs(a-e).max()/np.abs(e).max() for a,e in zip(ap,ex)]
        print(kind, l, P, "B",B,"p",p, " rel err (k, dk/dd, dk/dl): %.1e %.1e %.1e" % tuple(errs))
print("---- beta sweep")
for kind, l, P in [("eq",0.1,None),("eq",0.02,None),("per",1.1,0.55),("per",1.9,0.95),("per",0.4,1.5),("per",0.4,1.2),("per",3.0,0.3),("per",0.2,1.0)]:
    leff = l if kind=="eq" else P*l/(2*np.pi)
    ex = [ (c[:,None]*f).sum(0) for f in kern(kind, t[:,None]-z[None,:], l, P)]
    row=[]
    for B in (4,6,8,10,12,14,16,20,24,32,48,64):
        tn = nodes(0,T,B,16); m = moments(t,c,0,T,B,16)
        ap = [ (m[:,None]*f).sum(0) for f in kern(kind, tn[:,None]-z[None,:], l, P)]
        err = max(np.abs(a-e).max()/np.abs(e).max() for a,e in zip(ap,ex))
        row.append("B%d b=%.2f %.0e" % (B, (T/(2*B))/leff, err))
    print(kind,l,P,"leff=%.3f"%leff, " | ".join(row))
