#!/bin/bash
# FASTA reader: host encoder vs device encoder (row f-3) on a 1M x 500 bp file (~510 MB), wall clock of the SequenceSet constructor
mkdir -p gpurun_out
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
from bammmotif2_b200 import synth
fwd, sites, _ = synth.planted_sequences(9, ${NSEQ:-1000000}, 500, 12)
synth.write_fasta("$D/big.fasta", fwd)
PY
ls -la $D/big.fasta | awk '{print $5, "bytes"}' | tee gpurun_out/${1:-fasta}.txt
for m in 0 1 0 1; do
  echo -n "BAMM_DEVICE_FASTA=$m: " | tee -a gpurun_out/${1:-fasta}.txt
  BAMM_DEVICE_FASTA=$m bammmotif2_b200/bin/host_check parse STANDARD $D/big.fasta 0 device | tee -a gpurun_out/${1:-fasta}.txt
done
BAMM_TRACE=1 BAMM_DEVICE_FASTA=1 bammmotif2_b200/bin/host_check parse STANDARD $D/big.fasta 0 2>&1 | grep trace | tee -a gpurun_out/${1:-fasta}.txt
rm -rf $D
