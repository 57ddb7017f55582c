#!/bin/bash
# SASS of the hot kernels (from the object the library is linked from) -> profiles/r02_sass_*.txt
set -e
OBJ=build/obj/poa_kernels.cu.o
for spec in "2:convex:align" "1:affine:align" "2:convex:worker"; do
  g=$(echo $spec | cut -d: -f1); n=$(echo $spec | cut -d: -f2); k=$(echo $spec | cut -d: -f3)
  if [ $k = align ]; then
    f="_Z26poa_chain_align_kernel_p16ILi${g}ELb0EEvPK12PoaChainSlotPKiPK12PoaParamsDeviiii9P16Consts"
    out=profiles/r02_sass_poa_chain_align_kernel_p16_${n}.txt
    what="poa_chain_align_kernel_p16<GAP=${n}, TMA=false> (round schedule)"
  else
    f="_Z26poa_chain_dp_worker_kernelILi${g}ELb0EEvP12PoaChainSlotP12PoaChainSyncPK12PoaParamsDeviiii9P16Consts"
    out=profiles/r02_sass_poa_chain_dp_worker_kernel_${n}.txt
    what="poa_chain_dp_worker_kernel<GAP=${n}, TMA=false> (free-running schedule: wait loop + the same job function)"
  fi
  { echo "# cuobjdump -sass -fun '$f' $OBJ   (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)";
    echo "# $what: packed int16x2 forward DP (p16_run_job<GAP, GLOBAL, LEAN=true>) + backtrace, one warp per alignment";
    cuobjdump -res-usage $OBJ 2>/dev/null | grep -A1 "$f" | tail -1 | sed 's/^/# /';
    cuobjdump -sass -fun "$f" $OBJ | awk '/\/\*[0-9a-f]{4}\*\//{print}' | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1  /; s/\s+\/\*.*$//'; } > $out
  echo "$out: $(wc -l < $out) lines; DPX: $(grep -c 'VIMNMX\|VIADDMNMX\|VIADD.16x2' $out), SHFL: $(grep -c SHFL $out), LDS: $(grep -c 'LDS' $out), STS: $(grep -c 'STS' $out), LDG/LD.E: $(grep -c 'LD\.E' $out), ST.E: $(grep -c 'ST\.E' $out), spills (LDL/STL): $(grep -c 'LDL\|STL' $out)"
done
