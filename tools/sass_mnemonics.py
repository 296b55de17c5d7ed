"""cuobjdump -sass recommendation_b200/libgcf.so | python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.md
Counts, per kernel, the static SASS instructions that prove which data path a kernel uses (B200_PROFILING.md)."""
import collections
import re
import subprocess
import sys

cur, cnt = None, collections.OrderedDict()
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG[\.\w]*|UBLKCP[\.\w]*|SYNCS[\.\w]*|REDG[\.\w]*|RED[\.\w]*|LDG[\.\w]*|LDS[\.\w]*)\b")
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for mm in pat.findall(line):
        base = mm.split(".")[0]
        key = ".".join(mm.split(".")[:2]) if base == "UTMALDG" else base
        cnt[cur][key] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonics per kernel of the in-tree `libgcf.so` (`cuobjdump -sass`, static instruction counts)\n")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier.  They appear in the InfoNCE "
      "kernels and in the k-means assignment kernel (dense contractions).  The gather / scatter kernels are LDG / REDG by design: "
      "`tools/lab/gather_lab.cu` (`profiles/r02_gather_lab.log`) measures cp.async.bulk and TMA gather4 rings against LDG.128 on the "
      "same random 256-byte-row stream -- all three saturate the HBM copy rate, the async rings lose the L1 hits on hub rows.\n")
print("| kernel | UTCHMMA | LDTM | UTMALDG.2D | SYNCS | LDG | LDS | REDG+RED |\n|---|---|---|---|---|---|---|---|")
seen = set()
for n, d in zip(names, dem):
    c = cnt[n]
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("gcf::", "")
    tensor = any(k in c for k in ("UTCHMMA", "LDTM", "UTMALDG.2D"))
    if not tensor and not any(t in short for t in ("spmm_flat_kernel<16, 1, 8, false, false, 3", "spmm_flat_kernel<16, 1, 8, false, true, 3",
                                                   "spmm_csr_kernel<2, 1, 2, false, 4, 4, false", "spmm_csr_kernel<4, 1, 4, false, 4, 2, false",
                                                   "bpr_fused_kernel<16, 1, false, 4, 2", "gather_rows_kernel", "scatter_add_agg_kernel",
                                                   "peer_copy2d_kernel", "peer_sum_kernel", "peer_barrier_kernel")):
        continue
    if short in seen:
        continue
    seen.add(short)
    print(f"| `{short[:100]}` | {c.get('UTCHMMA', 0)} | {c.get('LDTM', 0)} | {c.get('UTMALDG.2D', 0)} | {c.get('SYNCS', 0)} | {c.get('LDG', 0)} | "
          f"{c.get('LDS', 0)} | {c.get('REDG', 0) + c.get('RED', 0)} |")
