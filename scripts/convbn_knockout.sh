# knock-out timing of the conv+BN unit (LG_CONVBN_DBG bits: 1 no A copies, 2 no MMAs, 4 no epilogue, 8 no B copies,
# 16 no proxy fences, 32 lane-0 polling); numbers with any bit set are timing probes, not results
for L in ${LAYERS:-Conv2d_2a_3x3 Mixed_6e.branch7x7dbl_2}; do
 for D in ${DBGS:-0 15 31 47 63 32}; do
  echo -n "dbg=$D  "; LG_CONVBN_DBG=$D timeout 60 python scripts/one_convbn.py $L 2>&1 | tail -1
 done
done
