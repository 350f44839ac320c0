"""Summarise an `ncu --page raw --csv` export: one row per profiled launch with the counters the
roofline needs (duration, DRAM bytes, tensor-pipe %, L2 throughput, registers)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "dur"),
        ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%el"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
        ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16ops%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("launch__registers_per_thread", "regs"),
        ("sm__cycles_elapsed.avg.per_second", "clk"),
        ]
print("\t".join(n for _, n in cols if _ in idx))
for r in data:
    out = []
    for h, n in cols:
        if h not in idx: continue
        v = r[idx[h]]
        if h == "Kernel Name": v = v[:28]
        u = units[idx[h]]
        out.append(f"{v}{(' '+u) if n in ('dur','rd','wr','clk') else ''}")
    print("\t".join(out))
