set -x
cd /root/repo
timeout 300 python tools/igev_feature_layout_probe.py > gpurun_out/r03f_igev_feature_layout.json 2> gpurun_out/r03f_probe.err; cat gpurun_out/r03f_igev_feature_layout.json; tail -3 gpurun_out/r03f_probe.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -k "gwc or conv3d or softargmin or deconv or hourglass or volume_stage_golden" > gpurun_out/r03f_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r03f_memcheck.log
