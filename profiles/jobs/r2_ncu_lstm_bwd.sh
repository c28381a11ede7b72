#!/bin/bash
# ncu --set full of the persistent LSTM backward kernel (one layer, T=20, Bq=4, H=1024)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_bwd_resident -s 3 -c 1 -f -o gpurun_out/lstm_bwd_resident \
   python profiles/lstm_bwd_time.py > gpurun_out/ncu_lstm_bwd.log 2>&1
python profiles/ncu_summary.py gpurun_out/lstm_bwd_resident.ncu-rep > gpurun_out/ncu_lstm_bwd_resident.txt 2>&1
python profiles/ncu_stalls.py gpurun_out/lstm_bwd_resident.ncu-rep 2>&1 | head -14 > gpurun_out/ncu_stalls_lstm_bwd_resident.txt
