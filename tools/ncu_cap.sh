# tools/ncu_cap.sh NAME KERNEL_REGEX OP [NAME KERNEL_REGEX OP ...] -- one `ncu --set full` capture per triple (release library, direct launches)
mkdir -p gpurun_out/ncu
while [ $# -ge 3 ]; do
  python tools/sweep.py --release --direct --ops $3 --steps 1 > gpurun_out/ncu/$1.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -o gpurun_out/ncu/$1 -f python tools/sweep.py --release --direct --ops $3 --steps 1 > gpurun_out/ncu/$1.ncu.log 2>&1
  shift 3
done
ls -la gpurun_out/ncu/*.ncu-rep
