"""Lists the mangled names of the reference's public operator surface as DEFINED by the reference's own objects
(oracle/_ref/obj/gs_*.o = /root/reference/gsplat/*.cpp compiled in place, fg_adam*.o, fg_rasterization_api.o from `make -C oracle fastgs_api_syms`) -> tests/golden/ref_symbols.txt.
Run in the build container after `make -C oracle ref`."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
want = re.compile(r"^_ZN6gsplat(23spherical_harmonics_fwd|23spherical_harmonics_bwd|14intersect_tile|16intersect_offset|"
                  r"16quats_to_rotmats|10relocation|9add_noise|24projection_ut_3dgs_fused|"
                  r"39rasterize_to_pixels_from_world_3dgs_fwd|39rasterize_to_pixels_from_world_3dgs_bwd)E")
names = set()
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    out = subprocess.run(["nm", "--defined-only", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    for line in out.splitlines():
        parts = line.split()
        if len(parts) == 3 and parts[1] in "TW":
            if want.match(parts[2]) or parts[2].startswith("_ZN7fast_gs9optimizer9adam_step") or \
                    parts[2].startswith("_ZN7fast_gs13rasterization15forward_wrapper") or \
                    parts[2].startswith("_ZN7fast_gs13rasterization16backward_wrapper"):
                names.add(parts[2])
path = os.path.join(ROOT, "tests", "golden", "ref_symbols.txt")
open(path, "w").write("\n".join(sorted(names)) + "\n")
print(len(names), "symbols ->", path)
