"""Turns an .ncu-rep (ncu --set full) into the markdown summary kept under profiles/.
Usage: python tools/ncu_summary.py report.ncu-rep "title" > profiles/xxx.md"""
import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
print(f"# {title}\n")
print(f"Source: `{rep}` (scratch; `ncu --set full --import-source on --clock-control none`, captured after the plain run exited 0).\n")
want = [("gpu__time_duration.sum", "duration (under ncu, cold caches, serialised)"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe (IMAD)"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (SHFL, LDS, LDG)"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor shared-memory read wavefronts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU shared-memory wavefronts (incl. SHFL)"),
        ("sm__inst_issued.avg.per_cycle_active", "warp instructions issued per cycle per SM (max 4)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
        ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / block"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 -> SM read sectors (32 B)"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
        ("smsp__inst_executed.sum", "warp instructions")]
for r in rows[2:]:
    d = dict(zip(h, r)); u = dict(zip(h, units))
    print(f"## {d.get('Kernel Name', '?').split('(')[0]}\n\n| metric | value |\n|---|---:|")
    for k, name in want:
        if d.get(k) not in (None, "", "n/a"):
            print(f"| {name} | {d[k]} {u.get(k, '')} |")
    st = sorted(((float(d[k]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in h
                 if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and d.get(k) not in (None, "", "n/a")), reverse=True)
    print("\n| stall reason (pc samples) | samples |\n|---|---:|")
    for v, k in st[:10]:
        print(f"| {k} | {int(v)} |")
    print()
