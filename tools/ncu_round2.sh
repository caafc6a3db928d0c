set -x
mkdir -p gpurun_out/ncu
bash tools/cli_trace.sh > gpurun_out/r2_cli_trace.txt 2>&1
cap() {  # name, kernel regex, ops
  python tools/sweep.py --release --direct --ops $3 --steps 1 > gpurun_out/ncu/$1.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -o gpurun_out/ncu/$1 -f python tools/sweep.py --release --direct --ops $3 --steps 1 > gpurun_out/ncu/$1.ncu.log 2>&1
}
cap geom_fliph4090 geom_kernel fliph_4090
cap geom_rot90_1080p geom_kernel rot90_1080p
cap conv_sep7 conv_sep_kernel gauss7
cap conv_dense7 conv_dp4a_kernel dense7
cap resize_rows imresize_rows16 resize_up
cap resize_cols imresize_colsK resize_up
cap conv3_16k conv3_strip conv3
python bench.py --steps 2 --warmup 3 --no-per-op --no-cpu --no-band > gpurun_out/ncu/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu/r2_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-per-op --no-cpu --no-band > gpurun_out/ncu/bench_ncu.log 2>&1
ls -la gpurun_out/ncu | head -40
