"""Summarise one kernel of an .ncu-rep (ncu --set full capture) as text: python scripts/ncu_summary.py rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(line for line in raw.splitlines() if line.startswith('"')))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = {h[i]: (v[i], u[i]) for i in range(len(h))}
        print("kernel:", d.get("Kernel Name", ("?",))[0])
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k][0]:>18s} {d[k][1]}")
        stalls = []
        for k, (val, _) in d.items():
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    stalls.append((float(val), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1.0
        print("  warp stall samples:", ", ".join(f"{n} {100 * s / tot:.1f}%" for s, n in sorted(stalls, reverse=True)[:8]))
        try:
            units = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            rd = float(d["dram__bytes_read.sum"][0]) * units[d["dram__bytes_read.sum"][1]]
            wr = float(d["dram__bytes_write.sum"][0]) * units[d["dram__bytes_write.sum"][1]]
            t = float(d["gpu__time_duration.sum"][0]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[d["gpu__time_duration.sum"][1]]
            print(f"  dram traffic {(rd + wr) / 1e9:.3f} GB in {t * 1e3:.3f} ms = {(rd + wr) / t / 1e9:.0f} GB/s (under ncu, cold, serialised)")
        except Exception:
            pass


def opcode_mix(rep, top=14):
    """executed warp-instructions by opcode (source page; needs -lineinfo / --import-source) and per output cell if given"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(line for line in raw.splitlines() if line.startswith('"')))
    if len(rows) < 3:
        return
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    if "Instructions Executed" not in ix:
        return
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    total = sum(int(r[ix["Instructions Executed"]]) for r in data)
    mix = {}
    for r in data:
        op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0]
        mix[op] = mix.get(op, 0) + int(r[ix["Instructions Executed"]])
    print(f"  executed warp-instructions {total}: " + ", ".join(f"{op} {100 * n / total:.1f}%" for op, n in sorted(mix.items(), key=lambda t: -t[1])[:top]))


if __name__ == "__main__":
    main(sys.argv[1])
    opcode_mix(sys.argv[1])
