#!/bin/bash
# Config 4 whole through the drop-in CLI: bin/BaMMmotif --EM --FDR (5 folds: EM to convergence on 4/5 of the positives, scoring of the
# held-out fifth and of the 10x sampled negatives, PR statistics, files) on NSEQ x 500 bp; wall clock with the stage trace.
mkdir -p gpurun_out
NSEQ=${NSEQ:-1000000}; TAG=${1:-c4full}; DEVS=${DEVS:-0}
D=$(mktemp -d)
python - <<PY
import sys, time; sys.path.insert(0, ".")
from bammmotif2_b200 import synth
t=time.time()
fwd, sites, _ = synth.planted_sequences(1234, $NSEQ, 500, 12)
synth.write_fasta("$D/in.fasta", fwd); synth.write_sites("$D/sites.block", sites)
print("FASTA written in %.1f s" % (time.time()-t))
PY
ls -la $D/in.fasta
ARGS="--bindingSiteFile $D/sites.block --EM -k 3 -K 2 --FDR -m 10 -n 5"
mkdir -p $D/ours
t0=$(date +%s.%N); BAMM_DEVICES=$DEVS BAMM_TRACE=1 timeout 1200 bammmotif2_b200/bin/BaMMmotif $D/ours $D/in.fasta $ARGS > $D/ours.log 2> $D/ours.err; rc=$?; t1=$(date +%s.%N)
echo "ours rc=$rc devices=$DEVS nseq=$NSEQ wall $(python -c "print('%.2f' % ($t1-$t0))") s" | tee gpurun_out/${TAG}.txt
grep "bamm host" $D/ours.err | tee -a gpurun_out/${TAG}.txt
tail -4 $D/ours.log | tee -a gpurun_out/${TAG}.txt; tail -3 $D/ours.err | tee -a gpurun_out/${TAG}.txt
ls -la $D/ours | tee -a gpurun_out/${TAG}.txt
head -3 $D/ours/*zoops.stats | tee -a gpurun_out/${TAG}.txt
nvidia-smi --query-gpu=memory.used --format=csv | tail -2
rm -rf $D
