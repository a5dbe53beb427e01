"""Summarise .ncu-rep captures (ncu --set full) into the text block kept under profiles/ (usage: ncu_summary.py rep...)."""
import csv, io, re, subprocess, sys
PAT = re.compile(r'^(gpu__time_duration\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|'
                 r'sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|'
                 r'launch__grid_size|launch__block_size|dram__bytes_read\.sum|dram__bytes_write\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
                 r'sm__inst_executed_pipe_(alu|lsu|xu|fma)\.avg\.pct_of_peak_sustained_active|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|'
                 r'smsp__inst_executed\.sum|sm__cycles_active\.avg|lts__t_sector_hit_rate\.pct)$')
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("----", r[hdr.index("Kernel Name")][:60], f"  [{rep.split('/')[-1]}]")
        stalls = []
        for h, u, v in zip(hdr, units, r):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
            elif PAT.match(h):
                print(f"  {h} = {v} {u}")
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"     {v:.3f} {h}")
