#!/bin/bash
# Round-2 evidence capture (run on the GPU box through gpurun).  The .ncu-rep files stay in /tmp on the box; what comes
# back in gpurun_out/ are the text summaries (profiles/ncu_summary.py, phase_shares.py, launch_shares.py).
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:train_bucket --launch-skip 1 -c 1 -o /tmp/bkt python profiles/train_only.py 48 2 > /tmp/a.log 2>&1
python profiles/ncu_summary.py /tmp/bkt.ncu-rep "K1b train_bucket_kernel<false,false>: one bench slab (69 120 points x 12 month groups, ref + hist), ncu --set full --clock-control none" > gpurun_out/r02_ncu_train_bucket.txt
ncu -i /tmp/bkt.ncu-rep --page source --csv --print-source sass > /tmp/bkt_src.csv 2>/dev/null
python profiles/phase_shares.py /tmp/bkt_src.csv >> gpurun_out/r02_ncu_train_bucket.txt
XSDBA_B200_TRAIN_ALGO=sort ncu --set full --clock-control none -k regex:train_fast --launch-skip 1 -c 1 -o /tmp/srt python profiles/train_only.py 48 2 > /tmp/a2.log 2>&1
python profiles/ncu_summary.py /tmp/srt.ncu-rep "K1f train_fast_kernel<false,false> (XSDBA_B200_TRAIN_ALGO=sort) on the same slab, same box" >> gpurun_out/r02_ncu_train_bucket.txt
ncu --set full --clock-control none -k regex:"adjust_tile|pack_tables|adjust_fix" --launch-skip 3 -c 3 -o /tmp/adj python profiles/train_only.py 48 2 adjust > /tmp/b.log 2>&1
python profiles/ncu_summary.py /tmp/adj.ncu-rep "adjust path of one bench slab: pack_tables + adjust_tile + adjust_fix" > gpurun_out/r02_ncu_adjust.txt
ncu --set full --clock-control none --import-source on -k regex:train_window -c 1 -o /tmp/win python profiles/cfg3_train_only.py 8 1 > /tmp/c.log 2>&1
python profiles/ncu_summary.py /tmp/win.ncu-rep "K1w train_window_kernel: config 3 train, 11 520 points x 365 day-of-year groups (window 31), nq = 100" > gpurun_out/r02_ncu_train_window.txt
ncu -i /tmp/win.ncu-rep --page source --csv --print-source sass > /tmp/win_src.csv 2>/dev/null
python profiles/phase_shares.py /tmp/win_src.csv >> gpurun_out/r02_ncu_train_window.txt
ncu --set full --clock-control none --import-source on -k regex:rank_window -c 1 -o /tmp/rkw python profiles/cfg3_train_only.py 8 1 adjust > /tmp/c2.log 2>&1
python profiles/ncu_summary.py /tmp/rkw.ncu-rep "K3w rank_window_kernel: config 3 adjust with rank_window=True, 11 520 points x 365 day-of-year groups (window 31), nq = 100" > gpurun_out/r02_ncu_rank_window.txt
ncu -i /tmp/rkw.ncu-rep --page source --csv --print-source sass > /tmp/rkw_src.csv 2>/dev/null
python profiles/phase_shares.py /tmp/rkw_src.csv >> gpurun_out/r02_ncu_rank_window.txt
# launch lists of the bench command, our kernels only (the synthetic-data generators of torch run outside the timed region)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"train_|pack_|adjust_|rank_" -c 400 --csv --log-file /tmp/l1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs none --e2e-rows 0 > /tmp/d.log 2>&1
python profiles/launch_shares.py /tmp/l1.csv > gpurun_out/r02_launches_bench_summary.csv
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"train_|pack_|adjust_|rank_|lookup" -c 400 --csv --log-file /tmp/l3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --configs cfg3 --e2e-rows 0 --lat-rows 48 > /tmp/e.log 2>&1
python profiles/launch_shares.py /tmp/l3.csv > gpurun_out/r02_launches_cfg3_summary.csv
ls -la gpurun_out
