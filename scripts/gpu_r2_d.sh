#!/bin/bash
# ncu full captures of the gen2 kernels (CUDA-core and DMMA) and of the gen1 DMMA variants; exported to CSV on the box
# (the reports with imported sources are 45 MB each, gpurun_out/ brings back 64 MiB)
mkdir -p gpurun_out/r2_d
cap() { tag=$1; gen=$2; mma=$3; pat=$4
  H3D_GEN2=$gen H3D_USE_MMA=$mma timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 8 -c 2 -o /tmp/$tag -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_d/ncu_$tag.log 2>&1; echo "$tag rc=$?"
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/r2_d/${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/$tag.ncu-rep --page source --csv > gpurun_out/r2_d/${tag}_source.csv 2>/dev/null
}
cap gen2_core 1 0 "k_volume2|k_gradient2"
cap gen2_mma 1 1 "k_volume2|k_gradient2"
cap gen1_mma 0 1 "k_volume|k_gradient"
ls -la gpurun_out/r2_d
