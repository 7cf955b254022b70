# ncu --set full of the gru08 z||r and q convs of the first loop iteration (conv_tc launches 69 and 70 of a bench step)
set -x
cd /root/repo
TAG=${1:-r03i}
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"conv_tc" -s 69 -c 2 -o gpurun_out/prof_iter_${TAG} -f python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_iter_${TAG}.log 2>&1
echo "rc=$?"; ls -la gpurun_out
