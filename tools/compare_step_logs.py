"""Compares the `full` ECM step logs of two tools/ecm_multi_gpu.py output files: step, (niter, nfeval) of each, relative
difference of the bound after the step.
    python tools/compare_step_logs.py a.json b.json"""
import json,sys
a=json.load(open(sys.argv[1])); b=json.load(open(sys.argv[2]))
for x,y in zip(a['full']['step_log'], b['full']['step_log']):
    print(x[1], x[3:], y[3:], abs(x[2]-y[2])/abs(y[2]))
