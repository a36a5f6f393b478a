"""Static evidence from the built library (no GPU needed): per-kernel SASS instruction mix and an excerpt of the densest
mma.m8n8k4.f64 (DMMA) region of the tensor-path kernels.
    python tools/sass_excerpt.py > profiles/r02_sass_mix_and_dmma_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "svgpfa_b200", "libsvgpfa_b200.so")
WANT = ("quad_latent_mma_kernelILi4ELb1ELi1", "quad_latent_mma_kernelILi4ELb0ELi1", "quad_latent_mma_kernelILi4ELb0ELi2",
        "quad_embed_cat_kernelILi5", "quad_embed_mma_kernelILi4",
        "indpoints_bwd_mma_kernelILi4", "panel_weights_kernelILi3ELi2", "panel_dC_kernelILi3", "panel_adjoint_kernelILb1",
        "panel_nodal_means_kernel", "panel_moments_kernelILi1", "kzz_chol_warp_kernelILb1", "indpoints_fwd_warp_kernel",
        "spike_tile_kernelILb1", "spike_gather_kernel", "lb_multidot_kernelILi3", "lb_combine_kernelILb0", "lb_update_kernel",
        "lb_step_kernelILb1", "lb_stats_kernel")
OPS = ("DMMA", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMG", "RED", "SHFL", "BAR",
       "IMAD", "LOP3", "IADD3", "BRA")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (sm_100a; {len(funcs)} kernels)\n")
    for want in WANT:
        for f in funcs:
            name = f.split("\n", 1)[0]
            if want not in name:
                continue
            lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
            ops = [re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l) for l in lines]
            ops = [m.group(1) for m in ops if m]
            mix = collections.Counter(o.split(".")[0] for o in ops)
            print(f"== {name}\n   {len(ops)} instructions; " + ", ".join(f"{k} {mix[k]}" for k in OPS if mix.get(k)))
            # densest 24-instruction window of DMMA
            idx = [i for i, o in enumerate(ops) if o.startswith("DMMA")]
            if idx:
                best = max(range(0, max(1, len(ops) - 24)), key=lambda s: sum(1 for i in idx if s <= i < s + 24))
                print("   densest DMMA window:")
                for l in lines[best:best + 24]:
                    code = re.sub(r"\s+", " ", l.split("*/", 1)[1].split("/*")[0]).strip()
                    print("      " + code)
            print()
            break


if __name__ == "__main__":
    main()
