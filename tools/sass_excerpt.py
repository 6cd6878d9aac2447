"""Mnemonic counts per kernel from `cuobjdump -sass` of the built library -> profiles/<tag>_sass_excerpt.md
usage: python tools/sass_excerpt.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "visual-odometry-project_b200", "libvo_b200.so")],
                     capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
want = ["harris_response_fast", "klt_track_packed", "klt_track_fast", "pyr_down_kernel", "harris_localmax", "harris_nms_scan", "harris_nms_bands",
        "p3p_solve_kernel", "p3p_count_kernel", "pipe_pose_kernel", "pipe_update_kernel", "pipe_regroup_kernel", "gftt_eig_kernel", "gftt_select_kernel",
        "knn2_kernel", "bgr2gray_kernel", "refine_pose_kernel", "triangulate_kernel", "bootstrap_kernel", "pipe_boot_apply_kernel"]
keys = ["UTMALDG", "SYNCS", "IDP.4A", "IDP.2A", "REDUX", "SHFL", "DFMA", "DMUL", "DADD", "MUFU", "I2F.F64", "LDS", "STS", "LDG", "STG", "ATOMS", "BAR",
        "VOTE", "POPC", "IMAD", "LOP3", "PRMT"]
out = [f"# SASS evidence (cuobjdump -sass libvo_b200.so, sm_100a), {tag}", "",
       "Mnemonic counts per kernel for the instructions the design relies on: TMA tile loads (UTMALDG) and mbarrier waits (SYNCS) in the",
       "Harris response kernel, integer dot products (IDP.4A = dp4a, IDP.2A = dp2a), warp-wide integer reductions (REDUX), shuffles, the",
       "unfused FP64 arithmetic of the geometry kernels (DMUL / DADD; DFMA only where the source asks for it).  Made by `tools/sass_excerpt.py`.", "",
       "| kernel | instrs | " + " | ".join(keys) + " |", "|---|---|" + "---|" * len(keys)]
samples = {}
for f in funcs:
    name = f.split("\n", 1)[0]
    short = next((w for w in want if w in name), None)
    if not short:
        continue
    lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
    ops = [m.group(1) for l in lines for m in [re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)] if m]
    c = collections.Counter()
    for o in ops:
        for k in keys:
            if o == k or o.startswith(k + "."):
                c[k] += 1
    t = re.findall(r"ILi(\d+)", name)
    label = short + (" <" + ", ".join(t) + ">" if t else "")
    out.append(f"| `{label}` | {len(ops)} | " + " | ".join(str(c[k]) for k in keys) + " |")
    if short in ("harris_response_fast", "klt_track_packed"):
        samples[label] = [l.strip() for l in lines if re.search(r"UTMALDG|SYNCS|IDP\.4A|IDP\.2A|REDUX", l)][:8]
out.append("")
for k, v in samples.items():
    out += [f"### `{k}`: first occurrences", "", "```"] + v + ["```", ""]
path = os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.md")
open(path, "w").write("\n".join(out))
print(path)
