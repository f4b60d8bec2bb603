"""Per-SASS-instruction view of a kernel's hot loop from an `ncu --set full --import-source on` report:
which opcodes the warp-state samples land on and why those warps were not issuing.
Usage: python profiles/scripts/hotloop_stalls.py gpurun_out/X.ncu-rep > profiles/X_hotloop.txt"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


ex = [f(r, "Instructions Executed") for r in data]
mx = max(ex)
hot = [i for i, e in enumerate(ex) if e > 0.9 * mx]
lo, hi = hot[0], hot[-1]
tot = sum(f(r, "# Samples") for r in data)
samp_hot = sum(f(r, "# Samples") for r in data[lo:hi + 1])
stall_keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("# %s" % rows[0][1])
print("# hot loop = the %d consecutive SASS instructions executed %.0f times per warp-loop (two timesteps per trip)" % (hi - lo + 1, mx))
print("# warp-state samples in the hot loop: %.1f %% of the kernel's %d" % (100 * samp_hot / tot, tot))
agg = collections.Counter()
byop, cnt, byop_stall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in data[lo:hi + 1]:
    op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]]).split()[0]
    s = f(r, "# Samples")
    byop[op] += s
    cnt[op] += 1
    for k in stall_keys:
        agg[k] += f(r, k)
        byop_stall[op][k] += f(r, k)
print("\n## why warps in the hot loop were not issuing (share of samples; `selected` = issuing)")
for k, v in agg.most_common(8):
    print("  %-24s %5.1f %%" % (k.replace("stall_", ""), 100 * v / samp_hot))
print("\n## per opcode: static count in the loop, share of the loop's samples, samples per instruction (loop average %.0f), top states" % (samp_hot / (hi - lo + 1)))
for op, v in byop.most_common(20):
    top = ", ".join("{} {:.0f}%".format(k.replace("stall_", ""), 100 * x / max(v, 1)) for k, x in byop_stall[op].most_common(3))
    print("  {:16s} n={:3d}  {:5.1f} %   {:5.0f} per instr   {}".format(op, cnt[op], 100 * v / samp_hot, v / cnt[op], top))
