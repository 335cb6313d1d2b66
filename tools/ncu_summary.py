#!/usr/bin/env python
"""Turn an `ncu --set full --import-source on` report into the two text artefacts kept under profiles/:

  python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r02_x        (run on the CPU box; needs `ncu`)

  <out>_ncu_raw.csv      selected raw metrics, one column per profiled launch (duration, DRAM bytes and
                         throughput, L2 throughput / hit rate, LTS->SM bytes, tensor-pipe activity, occupancy limits)
  <out>_hotspots.txt     warp-stall samples per SASS instruction: the top instructions with their stall reasons, the
                         tcgen05 / TMA / mbarrier instructions with execution counts (spin counts of the bounded
                         waits), and the mix of shared-memory access instructions (generic LD/ST vs LDS/STS)
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "launch__waves_per_multiprocessor", "launch__cluster_size"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(rep, out):
    rows = ncu_csv(rep, "raw")
    hdr, units, launches = rows[0], rows[1], rows[2:]
    with open(out + "_ncu_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        for m in ["ID"] + METRICS:
            if m in hdr:
                c = hdr.index(m)
                w.writerow([m, units[c]] + [r[c] for r in launches])
    print(f"{out}_ncu_raw.csv: {len(launches)} launches")


def hotspots(rep, out, top=30):
    rows = ncu_csv(rep, "source")
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    with open(out + "_hotspots.txt", "w") as f:
        for b in blocks:
            hdr, data = b["rows"][0], b["rows"][1:]
            if "# Samples" not in hdr:
                continue
            isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
            stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            total = sum(int(r[isamp]) for r in data if len(r) > isamp)
            key = (b["name"], total, len(data))
            if key in seen:                       # ncu repeats the block per section
                continue
            seen.add(key)
            f.write(f"== {b['name'][:110]}\n   {len(data)} SASS instructions, {total} warp-stall samples\n")
            agg = collections.Counter()
            for r in data:
                for c in stalls:
                    agg[hdr[c][6:]] += int(r[c])
            f.write("   stall reasons: " + ", ".join(f"{k} {v}" for k, v in agg.most_common(8) if v) + "\n\n")
            f.write("   top instructions by samples (index, samples, executed, SASS, main stalls)\n")
            for i in sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]:
                r = data[i]
                st = sorted(((hdr[c][6:], int(r[c])) for c in stalls if int(r[c])), key=lambda x: -x[1])[:3]
                f.write(f"   {i:5d} {int(r[isamp]):7d} {int(r[iex]):10d}  {r[isrc].strip()[:64]:64s} {st}\n")
            f.write("\n   async-machinery instructions (index, samples, executed): spins = executed of the in-loop TRYWAIT vs its first try\n")
            for i, r in enumerate(data):
                if any(k in r[isrc] for k in ("UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "SYNCS.PHASECHK", "BAR.SYNC", "UCGABAR")):
                    f.write(f"   {i:5d} {int(r[isamp]):7d} {int(r[iex]):10d}  {r[isrc].strip()[:80]}\n")
            mix = collections.Counter()
            for r in data:
                op = r[isrc].strip().split()
                op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
                for k in ("LDS", "STS", "LD.E", "ST.E", "LDG", "STG", "ATOM", "RED"):
                    if op.startswith(k):
                        mix[k] += 1
            f.write("\n   memory-instruction mix (static count): " + ", ".join(f"{k} {v}" for k, v in sorted(mix.items())) + "\n\n")
    print(f"{out}_hotspots.txt")


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    raw(sys.argv[1], sys.argv[2])
    hotspots(sys.argv[1], sys.argv[2])
