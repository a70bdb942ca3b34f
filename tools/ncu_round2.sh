#!/bin/bash
# Round-2 profiler evidence, run on a B200 box: bash tools/ncu_round2.sh  (outputs under gpurun_out/)
#  1. launch list of ONE steady-state training step with per-launch duration and DRAM bytes (single pass, no replay)
#  2. `ncu --set full` of the dominant kernels (one launch each), summarised with tools/ncu_keys.py
#  3. SASS mnemonic counts of libpnx.so (tcgen05 / TMA proof)
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --csv --log-file $O/launches_r2.csv python tools/ncu_step.py > $O/launches_r2.log 2>&1
python tools/launch_summary.py $O/launches_r2.csv > $O/launches_r2_summary.txt 2>&1

: > $O/ncu_r2_kernels.txt
cap() {  # name, kernel regex, script args...
  local name=$1 pat=$2; shift 2
  timeout 600 $NCU --set full --import-source on -k regex:$pat -c 1 "$@" > $O/ncu_$name.log 2>&1
}
cap win_fwd    igemm_win_kernel -o $O/r2_win_fwd    -f python tools/profile_gemm.py win_fwd
cap win_dgrad  igemm_win_kernel -o $O/r2_win_dgrad  -f python tools/profile_gemm.py win_dgrad
cap igemm_256  igemm_kernel     -o $O/r2_igemm_256  -f python tools/profile_gemm.py igemm_256
cap wgrad_head wgrad_kernel     -o $O/r2_wgrad_head -f python tools/profile_gemm.py wgrad_head
cap wgrad_256  wgrad_kernel     -o $O/r2_wgrad_256  -f python tools/profile_gemm.py wgrad_256
cap vox_mark   vox_mark_kernel  -o $O/r2_vox_mark   -f python tools/profile_vox.py 256
cap vox_rank   vox_rank_kernel  -o $O/r2_vox_rank   -f python tools/profile_vox.py 256
cap bn_bwd_apply  bn_bwd_apply_kernel  --profile-from-start off -s 8 -o $O/r2_bn_bwd_apply  -f python tools/ncu_step.py
cap bn_bwd_reduce bn_bwd_reduce_kernel --profile-from-start off -s 4 -o $O/r2_bn_bwd_reduce -f python tools/ncu_step.py
cap bn_apply      bn_apply_kernel      --profile-from-start off -s 30 -o $O/r2_bn_apply     -f python tools/ncu_step.py
cap pack          pack_weights_kernel  --profile-from-start off -o $O/r2_pack -f python tools/ncu_step.py
for n in win_fwd win_dgrad igemm_256 wgrad_head wgrad_256 vox_mark vox_rank bn_bwd_apply bn_bwd_reduce bn_apply pack; do
  echo "## $n" >> $O/ncu_r2_kernels.txt
  python tools/ncu_keys.py $O/r2_$n.ncu-rep "" >> $O/ncu_r2_kernels.txt 2>&1
  echo >> $O/ncu_r2_kernels.txt
done
cuobjdump -sass pillarnext_b200/libpnx.so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Za-z0-9_.]+" | awk '{print $2}' | \
  grep -E "^(UTCHMMA|UTCMMA|UTMALDG|UTMASTG|UTCBAR|UTCCP|SYNCS|UBLKCP|LDTM|STTM|REDG|RED|ATOMG|UTMAPF|UTMACCTL)" | sort | uniq -c | sort -rn > $O/sass_mnemonics_r2.txt
echo done
