"""Summarise an ncu report / launch list into profiles/*.md|csv (run here; no GPU needed)."""
import collections, csv, io, re, subprocess, sys

def launch_summary(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines); hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for row in r:
        if len(row) <= vi: continue
        v = float(row[vi].replace(",", ""))
        v = v / 1e3 if row[ui] == "ns" else (v * 1e3 if row[ui] == "ms" else v)
        name = re.sub(r"\(.*", "", row[ki])[:90]
        tot[name] = tot.get(name, 0) + v; cnt[name] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write('"%s",%d,%.1f,%.4f\n' % (k, cnt[k], v, v / total))
    return total

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
           "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

def rep_summary(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in METRICS or h in ("Kernel Name", "ID")]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (" [%s]" % units[i] if units[i] else "") for i in cols])
        for d in data:
            w.writerow([d[i][:80] for i in cols])

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launch_summary(sys.argv[2], sys.argv[3]))
    else:
        rep_summary(sys.argv[2], sys.argv[3])
