"""Key metrics of every kernel in an `ncu --set full` report (run where ncu is installed).
usage: python tools/ncu_summary.py report.ncu-rep [report2.ncu-rep ...] >> profiles/rN_ncu_*.txt"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TSCALE = {"s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"## {d['Kernel Name']}   grid={d.get('launch__grid_size')} block={d.get('launch__block_size')}   [{rep.split('/')[-1]}]")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k} [{u[k]}] = {d[k]}")
        try:
            b = float(d["dram__bytes_read.sum"]) * SCALE[u["dram__bytes_read.sum"]] + float(d["dram__bytes_write.sum"]) * SCALE[u["dram__bytes_write.sum"]]
            t = float(d["gpu__time_duration.sum"]) * TSCALE[u["gpu__time_duration.sum"]]
            print(f"  => dram traffic = {b / 1e6:.1f} MB per launch, {b / t / 1e9:.0f} GB/s under ncu")
        except Exception:
            pass
        print()
