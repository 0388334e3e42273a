"""Hot instructions of one kernel from an ncu report with source info: `python tools/ncu_hot.py report.ncu-rep [min_samples]`.
Prints the kernel's headline metrics and every SASS instruction with at least `min_samples` stall samples, with the CUDA
source line it was inlined into and the dominant stall reasons."""
import collections
import csv
import io
import subprocess
import sys


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    min_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
    for r in raw[2:]:
        print(r[idx["Kernel Name"]][:70], r[idx["Grid Size"]] if "Grid Size" in idx else "")
        for k in keys:
            if k in idx:
                print("   %-62s %s %s" % (k, r[idx[k]], units[idx[k]]))
    rows = page(rep, "source", ("--print-source", "cuda,sass"))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            sections.append(cur)
        elif r and r[0] == "Line No":
            cur["hdr"] = r
        elif r and r[0] == "Function Name":
            cur["fn"] = r[1]
        elif cur is not None and r:
            cur["rows"].append(r)
    # the last launch's sections: from the last occurrence of the first file name
    first = sections[0]["file"]
    start = max(i for i, s in enumerate(sections) if s["file"] == first)
    seen = {}
    for s in sections[start:]:
        h = s["hdr"]
        i_s = h.index("# Samples")
        stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        line = None
        for r in s["rows"]:
            if r[0].strip().isdigit():
                line = int(r[0])
            if not r[2]:
                continue
            try:
                addr = int(r[2], 16)
            except ValueError:
                continue
            n = int(r[i_s]) if r[i_s].isdigit() else 0
            st = collections.Counter()
            for i in stall:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0:
                    st[h[i][6:]] += v
            seen.setdefault(addr, []).append((s["file"].split("/")[-1], line, r[3], n, st))
    total = sum(v[0][3] for v in seen.values())
    print("total samples", total)
    for addr in sorted(seen):
        ent = seen[addr]
        n = ent[0][3]
        if n < min_samples:
            continue
        where = " <- ".join("%s:%s" % (f, ln) for f, ln, _, _, _ in ent)
        st = ent[0][4]
        print("%5d %4.1f%%  %-46s %-44s %s" % (n, 100.0 * n / total, ent[0][2][:46], where[:44],
                                              " ".join("%s=%d" % kv for kv in st.most_common(3))))


if __name__ == "__main__":
    main()
