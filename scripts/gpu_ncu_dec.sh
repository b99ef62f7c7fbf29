# ncu --set full of one decoder forward launch, for the default library and the variant libraries named in $VARIANTS
mkdir -p gpurun_out
for v in default ${VARIANTS}; do
  if [ $v = default ]; then unset UORF_B200_LIB; else export UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_$v.so; fi
  timeout 300 python bench.py --mode render --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$v.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_fwd -s 3 -c 1 -o gpurun_out/prof_dec_$v -f python bench.py --mode render --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dec_$v.log 2>&1
  tail -2 gpurun_out/ncu_dec_$v.log
done
ls -la gpurun_out/*.ncu-rep
