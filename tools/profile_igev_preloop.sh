# ncu passes over the IGEV pre-loop volume kernels (run on the GPU box through gpurun)
set -x
cd /root/repo
rm -f gpurun_out/*.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"conv3d_k|deconv3d|gwc_volume|softargmin" -c 20 --csv --log-file gpurun_out/r03d_igev_preloop_launches.csv python tools/igev_preloop_breakdown.py > gpurun_out/r03d_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"conv3d_k3" -c 3 -o gpurun_out/r03d_k3 python tools/igev_preloop_breakdown.py > gpurun_out/r03d_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"deconv3d" -s 2 -c 1 -o gpurun_out/r03d_deconv python tools/igev_preloop_breakdown.py > gpurun_out/r03d_ncu3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
