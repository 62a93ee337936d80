#!/bin/bash
# Drop-in CLI end to end at a size the reference still finishes: bin/BaMMmotif vs oracle/_ref/BaMMmotif_ref on the same
# synthetic FASTA, --EM --FDR (negative sampling + 5-fold cross-validation + PR statistics), wall clock and output diff.
mkdir -p gpurun_out
NSEQ=${NSEQ:-5000}; L0=${L0:-200}; TAG=${1:-cli}
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
from bammmotif2_b200 import synth
fwd, sites, _ = synth.planted_sequences(77, $NSEQ, $L0, 12)
synth.write_fasta("$D/in.fasta", fwd); synth.write_sites("$D/sites.block", sites)
PY
ARGS="--bindingSiteFile $D/sites.block --EM -k 2 -K 2 --FDR -m 10 -n 5"
mkdir -p $D/ours $D/ref
echo "input: $NSEQ x $L0 bp, args: --EM -k 2 -K 2 --FDR -m 10 -n 5" | tee gpurun_out/${TAG}_cli.txt
t0=$(date +%s.%N); bammmotif2_b200/bin/BaMMmotif $D/ours $D/in.fasta $ARGS > $D/ours.log 2>&1; rc=$?; t1=$(date +%s.%N)
echo "ours      rc=$rc wall $(python -c "print('%.2f' % ($t1-$t0))") s" | tee -a gpurun_out/${TAG}_cli.txt
tail -3 $D/ours.log | tee -a gpurun_out/${TAG}_cli.txt
t0=$(date +%s.%N); oracle/_ref/BaMMmotif_ref $D/ref $D/in.fasta $ARGS --threads $(nproc) > $D/ref.log 2>&1; rc=$?; t1=$(date +%s.%N)
echo "reference rc=$rc wall $(python -c "print('%.2f' % ($t1-$t0))") s ($(nproc) threads)" | tee -a gpurun_out/${TAG}_cli.txt
ls $D/ours $D/ref | tee -a gpurun_out/${TAG}_cli.txt
python - <<PY | tee -a gpurun_out/${TAG}_cli.txt
import numpy as np, glob, os
def table(p):
    rows = [l.split() for l in open(p) if l.strip() and not l.startswith("#")]
    try: return np.array([[float(x) for x in r] for r in rows if all(t.replace('.','',1).replace('e-','',1).replace('e+','',1).replace('-','',1).isdigit() for t in r)], float)
    except Exception: return None
for f in sorted(os.listdir("$D/ref")):
    a, b = os.path.join("$D/ref", f), os.path.join("$D/ours", f)
    if not os.path.exists(b): print("missing in ours:", f); continue
    same = open(a,"rb").read() == open(b,"rb").read()
    msg = "identical" if same else "differs"
    if not same and f.endswith((".ihbcp", ".ihbp", ".hbcp", ".hbp")):
        x = np.array([float(t) for l in open(a) for t in l.split()]); y = np.array([float(t) for l in open(b) for t in l.split()])
        msg += " (max rel diff %.2e over %d numbers, printed with 3-4 significant digits)" % (np.max(np.abs(x-y)/np.maximum(np.abs(x),1e-30)), len(x)) if len(x)==len(y) else " (different length)"
    print("%-40s %s" % (f, msg))
PY
rm -rf $D
