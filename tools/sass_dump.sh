#!/bin/bash
# Full SASS listings of the hot hand-written kernels, one file per kernel, from the in-tree build (runs WITHOUT a GPU):
#   tools/sass_dump.sh            -> profiles/sass/<kernel>.sass  (+ profiles/sass/INDEX.md with the mnemonics that prove
#                                    tcgen05 / TMEM / TMA / multimem: UTC*MMA, LDTM, UTMALDG, UTMASTG, UTMAREDG, LDGMC / STGMC / RED)
cd "$(dirname "$0")/.."
SO=federated_pytorch_test_b200/_build/fedb200_cuda/fedb200_cuda.so
OUT=profiles/sass
mkdir -p $OUT
rm -f $OUT/*.sass
declare -A KERNELS=(
  [igemm_persistent_128x3x2x1]='_ZN7fedb20023igemm_persistent_kernelILi128ELi3ELi2ELi1EEEv14CUtensorMap_stS1_S1_NS_11IgemmParamsE'
  [igemm_persistent_64x4x2x1]='_ZN7fedb20023igemm_persistent_kernelILi64ELi4ELi2ELi1EEEv14CUtensorMap_stS1_S1_NS_11IgemmParamsE'
  [conv3x3_ws_64]='_ZN7fedb20017conv3x3_ws_kernelILi64ELi30720ELi2EEEv14CUtensorMap_stS1_NS_10HaloParamsEii'
  [wgrad_tf32]='_ZN7fedb20017wgrad_tf32_kernelE14CUtensorMap_stS0_S0_NS_11WgradParamsE'
  [block_reduce]='_ZN7fedb20019block_reduce_kernelENS_8CommArgsE'
  [bb_update]='_ZN7fedb20016bb_update_kernelENS_6BBArgsE'
  [lbfgs_two_loop]='_ZN7fedb20021lbfgs_two_loop_kernelEPKfS1_PKiiiiS1_fPfS4_'
  [adam_prox]='_ZN7fedb20016adam_prox_kernelEPfPKfS0_S0_PKiiffffS2_S2_fffS2_'
  [bn_elu_fwd]='_ZN7fedb20017bn_elu_fwd_kernelEPKfPfS1_S1_S1_S2_S2_S2_S2_S2_iiffii'
  [bn_elu_bwd_reduce_1]='_ZN7fedb20024bn_elu_bwd_reduce_kernelILi1EEEvPKfS2_S2_S2_S2_S2_S2_Pfii'
  [bn_elu_bwd_apply_1]='_ZN7fedb20023bn_elu_bwd_apply_kernelILi1EEEvPKfS2_S2_S2_S2_S2_S2_PfS3_S3_S3_S3_iii'
  [convT_pack]='_ZN7fedb20017convT_pack_kernelEPKfPfiixxxx'
  [igemm_persistent_32x4x2x1]='_ZN7fedb20023igemm_persistent_kernelILi32ELi4ELi2ELi1EEEv14CUtensorMap_stS1_S1_NS_11IgemmParamsE'
  [gemm_f32]='_ZN7fedb20015gemm_f32_kernelILi32ELi32EEEvPKfS2_S2_Pfiiixxxxiii'
  [info_nce_fwd]='_ZN7fedb20019info_nce_fwd_kernelEPKfS1_iiPfS2_S2_'
)
echo "# SASS listings (cuobjdump -sass, sm_100a, from $SO)" > $OUT/INDEX.md
echo >> $OUT/INDEX.md
echo "| kernel | file | instructions | UTC*MMA | LDTM | UTMALDG | UTMASTG/REDG | multimem (LDGMC/STGMC/.MMA) | RED / ATOM |" >> $OUT/INDEX.md
echo "|---|---|---|---|---|---|---|---|---|" >> $OUT/INDEX.md
for k in "${!KERNELS[@]}"; do
  f=$OUT/$k.sass
  cuobjdump -sass -fun "${KERNELS[$k]}" $SO > $f 2>/dev/null
  n=$(grep -cE '^\s+/\*[0-9a-f]{4}\*/' $f)
  if [ "$n" = "0" ]; then echo "missing: $k" >&2; rm -f $f; continue; fi
  c() { grep -cE "$1" $f; }
  echo "| $k | $k.sass | $n | $(c 'UTC[A-Z]*MMA') | $(c 'LDTM') | $(c 'UTMALDG') | $(c 'UTMASTG|UTMAREDG') | $(c 'LDGMC|STGMC|MULTIMEM|\.MMA\.') | $(c ' RED|ATOM') |" >> $OUT/INDEX.md
done
sort -o $OUT/INDEX.md.tmp $OUT/INDEX.md 2>/dev/null; rm -f $OUT/INDEX.md.tmp
cat $OUT/INDEX.md
