#!/bin/bash
# Runs on the GPU box: the drop-in CLI alone on a synthetic FASTA with BAMM_TRACE=1 (stage times of the driver and of the library).
mkdir -p gpurun_out
NSEQ=${NSEQ:-30000}; L0=${L0:-200}; TAG=${1:-clitrace}
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
from bammmotif2_b200 import synth
fwd, sites, _ = synth.planted_sequences(77, $NSEQ, $L0, 12)
synth.write_fasta("$D/in.fasta", fwd); synth.write_sites("$D/sites.block", sites)
PY
mkdir -p $D/ours
t0=$(date +%s.%N)
BAMM_TRACE=1 bammmotif2_b200/bin/BaMMmotif $D/ours $D/in.fasta --bindingSiteFile $D/sites.block --EM -k 2 -K 2 --FDR -m 10 -n 5 > $D/ours.log 2> gpurun_out/${TAG}.err
t1=$(date +%s.%N)
echo "wall $(python -c "print('%.2f' % ($t1-$t0))") s; launched at $t0, reaped at $t1" | tee gpurun_out/${TAG}.txt
grep "bamm host" gpurun_out/${TAG}.err | tee -a gpurun_out/${TAG}.txt
grep -c "bamm trace" gpurun_out/${TAG}.err
tail -2 $D/ours.log
rm -rf $D
