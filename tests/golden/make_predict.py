"""Generates tests/golden/predict/*.npz: the post-fit read-outs of the UNMODIFIED reference on a time grid that is NOT
the quadrature grid:
    predictLatents               stats/svLowerBound.py:116-117 -> stats/svPosteriorOnLatents.py:57-77
    predictEmbedding             stats/svLowerBound.py:119-120 -> stats/svEmbedding.py:86-92
    computeExpectedPosteriorCIFs stats/svLowerBound.py:64-66   -> stats/expectedLogLikelihood.py:62-73
for inputs that are committed golden fixtures (the fixture names the case it belongs to).

    python tests/golden/make_predict.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from svgpfa_b200 import synthetic  # noqa: E402
import ref_harness  # noqa: E402


def main():
    torch.set_num_threads(8)
    os.makedirs(os.path.join(HERE, "predict"), exist_ok=True)
    for name, T, lo, hi in (("tiny_mixed", 37, -0.05, 1.07), ("matlab_r5", 53, 0.3, 19.9), ("config3_r4", 41, 0.0, 1.0)):
        case, _ = synthetic.load_case(os.path.join(HERE, name + ".npz"))
        R = case["spike_counts"].shape[0]
        rng = np.random.default_rng(17)
        # a different irregular grid per trial, unsorted, including points outside the inducing-point range
        times = np.stack([rng.permutation(np.linspace(lo, hi, T) + rng.uniform(-0.01, 0.01, T)) for _ in range(R)])[:, :, None]
        model, _ = ref_harness.build_reference_model(case, requires_grad=False)
        model.buildKernelsMatrices()
        tt = torch.from_numpy(times)
        with torch.no_grad():
            mu, var = model.predictLatents(times=tt)
            e_mu, e_var = model.predictEmbedding(times=tt)
            cifs = model.computeExpectedPosteriorCIFs(times=tt)
        cif = np.stack([np.stack([c.numpy() for c in trial], axis=1) for trial in cifs])      # (R, T, N)
        path = os.path.join(HERE, "predict", name + ".npz")
        np.savez_compressed(path, times=times, latent_mean=mu.numpy(), latent_var=var.numpy(),
                            embedding_mean=e_mu.numpy(), embedding_var=e_var.numpy(), cif=cif)
        print(name, times.shape, mu.shape, e_mu.shape, cif.shape, f"{os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    main()
