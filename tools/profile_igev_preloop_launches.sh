cd /root/repo
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"conv3d_k|deconv3d|gwc_volume|softargmin" -c 20 --csv --log-file gpurun_out/r03r_igev_preloop_launches.csv python tools/igev_preloop_breakdown.py > gpurun_out/r03r_ncu1.log 2>&1; echo "ncu1 rc=$?"
