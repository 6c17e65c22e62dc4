#!/bin/bash
# SASS opcode evidence of the built library (no GPU needed): scratch/sass_counts.sh > profiles/r02_sass_counts.txt
LIB=${1:-rpsmf_b200/libpsmf_b200.so}
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo "# cuobjdump -sass $LIB  ($(stat -c %s "$LIB") bytes, $(grep -c 'Function :' "$TMP") kernels, arch $(grep -m1 -o 'sm_[0-9a]*' "$TMP"))"
echo "# opcode                what it proves                                                              count"
for pat in "UBLKCP:cp.async.bulk (1-D TMA bulk copies HBM<->shared memory: chunk ring, batch kernel)" \
           "SYNCS:mbarrier arrive / try_wait (completion of the bulk copies, warp hand-offs)" \
           "DMMA:fp64 tensor-core MMA m8n8k4 (masked Gram of a 32-row tile)" \
           "DFMA:fp64 FMA (rank-1 update, y_hat, r x r algebra)" \
           "REDUX:warp-wide integer max (pivot search of the Gauss-Jordan elimination)" \
           "NANOSLEEP:polling back-off" \
           "MUFU.RCP64H:reciprocal seed of fast_rcp (pivots)" \
           "UTMALDG:tensor-map TMA loads (not used: the tiled C makes every chunk one contiguous block)" \
           "UTCHMMA\|UTCQMMA\|UTCOMMA:tcgen05 MMA (not used: no fp64 on tcgen05, r <= 16)" \
           "LDTM:TMEM loads (not used)" \
           "STL:local-memory stores (spills)" \
           "LDL:local-memory loads (spills)"; do
  op=${pat%%:*}; what=${pat#*:}
  printf "%-22s %-75s %s\n" "${op//\\|/|}" "$what" "$(grep -c "$op" "$TMP")"
done
echo "# per kernel family (r = 16, double)"
for k in psmf_stream_kernelILi16Ed psmf_filter_kernelILi16EdLb0 psmf_filter_kernelILi16EdLb1 psmf_batch_kernelILi8EdLi4; do
  awk -v k="$k" '/Function :/{on = index($0, k) > 0} on' "$TMP" > "$TMP.k"
  printf "%-40s UBLKCP %s  SYNCS %s  DMMA %s  DFMA %s  REDUX %s  STL %s  LDL %s  BAR %s\n" "$k" "$(grep -c UBLKCP $TMP.k)" "$(grep -c SYNCS $TMP.k)" \
    "$(grep -c DMMA $TMP.k)" "$(grep -c DFMA $TMP.k)" "$(grep -c REDUX $TMP.k)" "$(grep -c 'STL' $TMP.k)" "$(grep -c 'LDL' $TMP.k)" "$(grep -c 'BAR\.' $TMP.k)"
done
rm -f "$TMP" "$TMP.k"
