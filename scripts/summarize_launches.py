"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py:
per-kernel share of ONE apply+factor+solve step (the launches between two
consecutive leaf-apply kernels), as a markdown table.
usage: python scripts/summarize_launches.py gpurun_out/r1c_launches.csv [step_index] > profiles/...md"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    ui = hdr.index("Metric Unit")
    out = []
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "").replace("sb200::<unnamed>::", "")
        out.append((int(r[ii]), name, v * scale))
    return out


def main():
    path = sys.argv[1]
    L = load(path)
    ends = [i for i, (_, n, _) in enumerate(L) if n.startswith("hss_leaf_kernel") or n.startswith("hss_leaf_mm_kernel")]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else len(ends) // 2
    # a step = [first up-sweep launch after the previous step's last solve kernel ... last bwd launch]
    # the leaf-apply kernel closes the apply; the step continues through factor and solve until the
    # next step's first hss_up launch
    lo = ends[k - 1] + 1
    while not L[lo][1].startswith("ulv_"):
        lo += 1
    # walk back: the apply of step k starts after the last ulv_bwd of step k-1
    prev_end = max(i for i in range(ends[k - 1]) if L[i][1].startswith("ulv_bwd")) if k >= 2 else -1
    first = prev_end + 1
    last = max(i for i in range(ends[k - 1], ends[k]) if L[i][1].startswith("ulv_bwd"))
    step = L[first:last + 1]
    agg = collections.OrderedDict()
    for _, n, ms in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"# launch list of one step (launch ids {step[0][0]}..{step[-1][0]}, {len(step)} launches, "
          f"sum of device times {tot:.3f} ms; cold-cache, serialised under ncu: compare SHARES)\n")
    print(f"source: `{path}`\n")
    print("| kernel | launches | ms | share |")
    print("|---|---|---|---|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{n}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % |")


if __name__ == "__main__":
    main()
