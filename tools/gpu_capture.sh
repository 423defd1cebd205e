# usage (on the GPU box): bash tools/gpu_capture.sh <tag> [tests] [ncu]
#   default bench line -> gpurun_out/<tag>_bench.json; optional -m gpu tests; optional ncu launch list + --set full capture
TAG=${1:-run}
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err | cut -c1-400
python tools/show_bench.py gpurun_out/${TAG}_bench.json
if [ "$3" = "ncu" ]; then
  # launch list of the default workload's step (bench exited 0 above): skip the warm-up launches by name filter later
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_secondary --no_parity --no_kernel_pass > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'episode_(fwd|bwd)_v2|gram_tc|gemm_x3' -s 12 -c 6 -o gpurun_out/${TAG}_full -f \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_secondary --no_parity --no_kernel_pass > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  ls -la gpurun_out/${TAG}_full.ncu-rep
fi
