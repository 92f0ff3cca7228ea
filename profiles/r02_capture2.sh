set -x
ncu --set full --clock-control none --import-source on -k regex:train_bucket --launch-skip 1 -c 1 -o /tmp/bkt python profiles/train_only.py 48 2 > /tmp/a.log 2>&1
python profiles/ncu_summary.py /tmp/bkt.ncu-rep "K1b" > gpurun_out/n_bkt.txt
ncu -i /tmp/bkt.ncu-rep --page source --csv --print-source sass > /tmp/bkt_src.csv 2>/dev/null
python profiles/phase_shares.py /tmp/bkt_src.csv >> gpurun_out/n_bkt.txt
gzip -c /tmp/bkt_src.csv > gpurun_out/bkt_src.csv.gz
ncu --set full --clock-control none --import-source on -k regex:"adjust_tile|pack_tables|adjust_fix" --launch-skip 3 -c 3 -o /tmp/adj python profiles/train_only.py 48 2 adjust > /tmp/b.log 2>&1
python profiles/ncu_summary.py /tmp/adj.ncu-rep "adjust" > gpurun_out/n_adj.txt
ncu -i /tmp/adj.ncu-rep --page source --csv --print-source sass -k regex:adjust_tile > /tmp/adj_src.csv 2>/dev/null
python profiles/phase_shares.py /tmp/adj_src.csv >> gpurun_out/n_adj.txt
gzip -c /tmp/adj_src.csv > gpurun_out/adj_src.csv.gz
ls -la gpurun_out
