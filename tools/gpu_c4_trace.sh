#!/bin/bash
# where the time of the whole --FDR run (config 4, 1M positives) goes inside the library: BAMM_TRACE lines summed by call and phase
mkdir -p gpurun_out
NSEQ=${NSEQ:-1000000}; TAG=${1:-c4trace}
D=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
from bammmotif2_b200 import synth
fwd, sites, _ = synth.planted_sequences(1234, $NSEQ, 500, 12)
synth.write_fasta("$D/in.fasta", fwd); synth.write_sites("$D/sites.block", sites)
PY
mkdir -p $D/ours
BAMM_TRACE=1 bammmotif2_b200/bin/BaMMmotif $D/ours $D/in.fasta --bindingSiteFile $D/sites.block --EM -k 3 -K 2 --FDR -m 10 -n 5 --verbose > $D/ours.log 2> $D/ours.err
grep -c " iter, llh=" $D/ours.log | sed 's/^/EM iterations printed: /' | tee gpurun_out/${TAG}.txt
python - <<PY | tee -a gpurun_out/${TAG}.txt
import collections, re
tot = collections.OrderedDict()
for l in open("$D/ours.err"):
    m = re.match(r"\[bamm trace\] (\S+): (.*) ([0-9.]+) ms", l)
    if m:
        k = m.group(1) + ": " + m.group(2)
        tot[k] = tot.get(k, 0.0) + float(m.group(3))
    elif l.startswith("[bamm host]"):
        print(l.strip())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
    print("%9.1f ms  %s" % (v, k))
PY
rm -rf $D
