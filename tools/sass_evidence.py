"""profiles/rNN_ptxas.txt + profiles/rNN_sass_histogram.txt: registers / spills / shared memory per kernel
(`nvcc -Xptxas -v`) and an instruction-mix histogram of the SHIPPED library's SASS (`cuobjdump -sass`).

    python tools/sass_evidence.py r02
"""
import collections
import os
import re
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(R, "local-feature-refinement_b200", "csrc")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"

# ---- ptxas -v
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
       "-Xptxas", "-v", "-c", "-o", "/tmp/lfr_capi_evidence.o", os.path.join(CSRC, "lfr_capi.cu")]
err = subprocess.run(cmd, capture_output=True, text=True).stderr
rows = []
cur = None
for line in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur and not any(r[0] == cur for r in rows):
        rows.append([cur, m.group(1), m.group(2), m.group(3), "", ""])
        continue
    m = re.search(r"Used (\d+) registers, used (\d+) barriers", line)
    if m and cur:
        for r in rows:
            if r[0] == cur and not r[4]:
                r[4], r[5] = m.group(1), m.group(2)
with open(os.path.join(R, "profiles", "%s_ptxas.txt" % tag), "w") as f:
    f.write("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v  csrc/lfr_capi.cu\n")
    f.write("%-58s %9s %12s %12s %6s %8s\n" % ("kernel", "stack_B", "spill_st_B", "spill_ld_B", "regs", "barriers"))
    for r in rows:
        f.write("%-58s %9s %12s %12s %6s %8s\n" % tuple(r))

# ---- SASS histogram of the shipped library
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, "liblfr_b200.so")], capture_output=True, text=True).stdout
hist = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = re.sub(r"\(.*", "", subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip())
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1).split(".")[0]
        hist[name][op] += 1
        hist[name]["_total"] += 1
        if m.group(1).startswith("LDG.E.128") or m.group(1).startswith("LDS.128") or m.group(1).startswith("STS.128"):
            hist[name][m.group(1).split(".")[0] + ".128"] += 1
cols = ["_total", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "LDS.128", "STS", "STS.128", "LDG", "LDG.128", "STG", "LDL", "STL",
        "UBLKCP", "SYNCS", "SHFL", "BAR", "WARPSYNC", "BRA", "CALL", "HMMA", "UTCHMMA"]
with open(os.path.join(R, "profiles", "%s_sass_histogram.txt" % tag), "w") as f:
    f.write("# cuobjdump -sass csrc/liblfr_b200.so (sm_100a): instruction counts per kernel; _total x 16 B = code size.\n"
            "# UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier ops; no HMMA / UTC*MMA: fp64 sparse NLLS, no tensor cores.\n")
    f.write("%-44s " % "kernel" + " ".join("%8s" % c for c in cols) + "\n")
    for k, h in hist.items():
        f.write("%-44s " % k[:44] + " ".join("%8d" % h.get(c, 0) for c in cols) + "\n")
print(open(os.path.join(R, "profiles", "%s_ptxas.txt" % tag)).read())
print(open(os.path.join(R, "profiles", "%s_sass_histogram.txt" % tag)).read())
