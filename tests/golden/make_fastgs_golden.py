"""Records golden vectors of the UNMODIFIED reference fastgs CUDA kernels (oracle/_ref/libfastgs_ref.so) for the
seeded case tests/gpu_diag.py:fastgs_case.  Run ON THE GPU BOX (the reference needs a GPU):
    gpurun -- 'python tests/golden/make_fastgs_golden.py gpurun_out/fastgs_ref_golden.npz'
then copy the file to tests/golden/.  tests/test_oracle_golden_fastgs.py pins the CPU oracle against it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_diag as D  # noqa: E402
import ref_libs as R  # noqa: E402

CASES = {"a": dict(n=600, w=96, h=80, deg=3, seed=11, sigma_px=4.0), "b": dict(n=400, w=64, h=64, deg=1, seed=12, sigma_px=6.0)}


def main(path):
    out = {}
    fg = R.FastGS()
    for name, kw in CASES.items():
        c = D.fastgs_case(**kw)
        t = {k: torch.as_tensor(np.ascontiguousarray(c[k], np.float32)).cuda() for k in (
            "means", "scales_raw", "rotations_raw", "opacities_raw", "sh0", "shN", "w2c", "cam_pos", "grad_image", "grad_alpha")}
        img, alpha, counts = fg.forward(t["means"], t["scales_raw"], t["rotations_raw"], t["opacities_raw"], t["sh0"],
                                        t["shN"], t["w2c"], t["cam_pos"], c["active_sh_bases"], c["width"], c["height"],
                                        c["fx"], c["fy"], c["cx"], c["cy"])
        g = fg.backward(t["grad_image"], t["grad_alpha"], img, alpha, t["means"], t["scales_raw"], t["rotations_raw"],
                        t["shN"], t["w2c"], t["cam_pos"], c["active_sh_bases"], c["width"], c["height"], c["fx"], c["fy"],
                        c["cx"], c["cy"])
        torch.cuda.synchronize()
        out[f"{name}_kw"] = np.array([kw["n"], kw["w"], kw["h"], kw["deg"], kw["seed"], int(kw["sigma_px"] * 1000)])
        out[f"{name}_counts"] = np.array(counts)
        out[f"{name}_image"], out[f"{name}_alpha"] = img.cpu().numpy(), alpha.cpu().numpy()
        for k, v in g.items():
            out[f"{name}_grad_{k}"] = v.cpu().numpy()
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "fastgs_ref_golden.npz"))
