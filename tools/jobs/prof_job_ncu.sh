mkdir -p gpurun_out/r02/ncu /tmp/ncu
M="--set full --clock-control none --import-source on"
cap() {  # name, kernel regex, command...
  name=$1; rex=$2; shift 2
  timeout 300 ncu $M -k regex:$rex -c 1 -s 1 -o /tmp/ncu/$name -f "$@" > gpurun_out/r02/ncu/$name.log 2>&1
  python tools/ncu_extract.py /tmp/ncu/$name.ncu-rep > gpurun_out/r02/ncu/ncu_full_$name.txt 2>&1
  rm -f /tmp/ncu/$name.ncu-rep
}
cap head_128_80 conv_tc python tools/run_head.py 128 80 64
cap conv_64_64_1_1_160_160 conv_tc python tools/run_layer.py 64 64 1 1 160 160 64 3
cap conv_128_128_1_1_40_40 conv_tc python tools/run_layer.py 128 128 1 1 40 40 64 3
cap conv_16_32_3_1_320_320_stem conv_tc python tools/run_layer.py 16 32 3 1 320 320 64 3
cap conv_128_128_3_1_40_40 conv_tc python tools/run_layer.py 128 128 3 1 40 40 64 3
cap tconv_64_64_3_80_80_b16 tconv python tools/run_train_kernels.py 64 64 3 80 80 16
cap twgrad_64_64_3_80_80_b16 twgrad python tools/run_train_kernels.py 64 64 3 80 80 16
head -8 gpurun_out/r02/ncu/ncu_full_*.txt | cut -c1-150
cat gpurun_out/r02/ncu/*.log | grep -v PROF | tail -12
