"""Shared helpers for the test-suite: golden-fixture loading and a FASTA reader that mirrors the
reference's (src/init/SequenceSet.cpp:67-225) closely enough for the fixtures."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["jund_k2", "syn_k2_N", "syn_ss_k1_q", "syn_k4", "syn_ext_k1", "syn_k3_fdr", "syn_k0"]


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = {}
        for line in bytes(self.z["meta"]).decode().splitlines():
            t = line.split()
            if t[0] == "motif":
                self.meta[t[2]] = int(t[3])
            else:
                self.meta[t[0]] = float(t[1]) if "." in t[1] or "e" in t[1] else int(t[1])
        self.args = bytes(self.z["args"]).decode().split()
        self.A = self.meta["A"]
        self.K = self.meta["K"]
        self.K_bg_model = self.meta["K_bg_model"]
        self.K_bg = min(self.K, self.K_bg_model)
        self.W = self.meta["W"]
        self.q = np.float32(self.meta["q"])
        self.ss = bool(self.meta["ss"])
        self.iterations = self.meta["iterations"]
        self.optimize_q = "--optimizeQ" in self.args
        self.alphabet = self.args[self.args.index("--alphabet") + 1] if "--alphabet" in self.args else "STANDARD"

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z.files

    def fasta_records(self):
        return parse_fasta(bytes(self.z["fasta_text"]).decode())

    def sites(self):
        return [l for l in bytes(self.z["sites_text"]).decode().split("\n") if l]

    def bg_alpha(self):
        a = np.full(self.K_bg_model + 1, 10.0, np.float32)   # reference: Global.cpp:48, 274-278
        a[0] = 1.0
        return a


def parse_fasta(text):
    recs, header, seq = [], None, []
    for line in text.split("\n"):
        if not line:
            continue
        if line[0] == ">":
            if header is not None and seq:
                recs.append((header, "".join(seq)))
            header = line.split("\t")[0].split("\r")[0] if len(line) > 1 else ">"
            seq = []
        else:
            seq.append(line)
    if header is not None and seq:
        recs.append((header, "".join(seq)))
    return recs


def encode_text(seq, base_to_code):
    return base_to_code[np.frombuffer(seq.encode(), np.uint8)]


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else np.finfo(np.float64).tiny)
