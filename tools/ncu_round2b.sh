set -x
mkdir -p gpurun_out/ncu
cap() {  # name, kernel regex, ops
  python tools/sweep.py --release --direct --ops $3 --steps 1 > gpurun_out/ncu/$1.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -o gpurun_out/ncu/$1 -f python tools/sweep.py --release --direct --ops $3 --steps 1 > gpurun_out/ncu/$1.ncu.log 2>&1
}
cap conv3_ua conv3_strip_ua conv3_odd
cap geom_rot90_4090 geom_kernel rot90_4090
cap rows_mono4090 rows_kernel mono_4090
python bench.py --steps 1 --warmup 3 --no-per-op --no-cpu --no-band --no-graph > gpurun_out/ncu/headline_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3_strip_kernel -s 600 -c 1 -o gpurun_out/ncu/conv3_headline_16k -f python bench.py --steps 1 --warmup 3 --no-per-op --no-cpu --no-band --no-graph > gpurun_out/ncu/headline_ncu.log 2>&1
ls -la gpurun_out/ncu/*.ncu-rep
