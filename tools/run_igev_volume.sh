set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "igev or gwc or conv3d or softargmin or deconv" > gpurun_out/r03c_pytest_igev.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r03c_pytest_igev.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv3d_k3|deconv3d|gwc_volume|conv3d_k1" -c 40 -o gpurun_out/r03c_igev_preloop python tools/igev_preloop_breakdown.py > gpurun_out/r03c_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
