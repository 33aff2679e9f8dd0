#!/bin/bash
# ncu evidence for round 2 (ONE GPU, under gpurun).  Numbers printed while running under ncu are never bench values.
#   gpurun --timeout 1500 -- 'bash tools/ncu_r02.sh'
# Outputs (gpurun_out/): ncu_r02_mega.ncu-rep (+ .csv raw page), launch lists of a batch-32 decode step and of a prefill,
# full captures of the batch-32 decode attention / stream-K GEMM and of the 256-token prefill tiles.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batch32 --no-extras --graph 0"
# 1. the persistent decode kernel: one launch = one whole decode step (Llama-2-7B, batch 1, ctx ~ 520)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel -s 24 -c 1 -f -o $O/ncu_r02_mega \
    $B > $O/ncu_r02_mega.log 2>&1
ncu -i $O/ncu_r02_mega.ncu-rep --page raw --csv > $O/ncu_r02_mega_raw.csv 2>/dev/null
# 2. launch list of batch-32 DECODE steps only (tc_gemm_sk / attn_decode / rmsnorm / argmax / embed are decode-side kernels;
#    prefill uses tc_gemm_kernel / attn_prefill): ~7 kernels x 32 layers per step
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm_sk_kernel|attn_decode_kernel|rmsnorm_kernel|argmax_kernel|embed_kernel' \
    -s 4000 -c 460 --csv --log-file $O/ncu_r02_launches_decode_b32.csv $B --batch 32 > /dev/null 2>&1
# 3. launch list of one prefill (512 tokens, batch 1): tc_gemm_kernel<128|256> + attn_prefill + norms
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm_kernel|attn_prefill_kernel|rmsnorm_kernel|proj_rows_kernel|argmax_kernel|embed_kernel' \
    -s 0 -c 240 --csv --log-file $O/ncu_r02_launches_prefill_b1.csv $B > /dev/null 2>&1
# 4. full captures: batch-32 decode attention (x2) and stream-K gate/up GEMM (x4)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 700 -c 2 -f -o $O/ncu_r02_attn_b32 $B --batch 32 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_sk_kernel -s 2800 -c 4 -f -o $O/ncu_r02_tcsk_b32 $B --batch 32 > /dev/null 2>&1
# 5. full capture: prefill GEMMs of a 32 x 512-token prefill (1024-row chunks -> 256-token tiles)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 8 -c 5 -f -o $O/ncu_r02_prefill_b32 $B --batch 32 > /dev/null 2>&1
for f in ncu_r02_attn_b32 ncu_r02_tcsk_b32 ncu_r02_prefill_b32; do
  [ -f $O/$f.ncu-rep ] && ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
ls -la $O/ncu_r02_* | cut -c30-
python - <<'PY'
import csv, glob
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_hmma.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warp_latency_issue_stalled_barrier.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for f in sorted(glob.glob("gpurun_out/ncu_r02_*_raw.csv")):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        print(f, "empty"); continue
    hdr = rows[0]
    print("==", f)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?")[:60]
        out = [name]
        for k in keys:
            if k in d and d[k] != "":
                out.append(f"{k.split('.')[0][-28:]}={d[k]}")
        print("  ", " | ".join(out)[:900])
PY
