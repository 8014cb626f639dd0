"""SASS evidence of the Blackwell-native instructions in libriders_b200.so: counts per kernel of the tcgen05 / TMEM / TMA
mnemonics (and the legacy mma.sync ones), with one example line each.  Usage: python tools/sass_evidence.py > profiles/..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "riders_b200", "libriders_b200.so")
PATS = [("UTCHMMA", r"\bUTCHMMA\b"), ("UTCQMMA", r"\bUTCQMMA\b"), ("UTCMMA", r"\bUTC[A-Z]*MMA\b"), ("UTCBAR", r"\bUTCBAR\b"),
        ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UBLKCP", r"\bUBLKCP\b"), ("UBLKPF", r"\bUBLKPF\b"),
        ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("SYNCS", r"\bSYNCS\b"), ("HMMA", r"\bHMMA\b"), ("LDSM", r"\bLDSM\b")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], check=True, capture_output=True, text=True).stdout
    print("# SASS evidence of the Blackwell-native instructions in riders_b200/libriders_b200.so (cuobjdump -sass, sm_100a),")
    print("# counted per kernel.  UTCHMMA = tcgen05.mma kind::f16 / tf32, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,")
    print("# UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch.L2, UTMALDG / UTMASTG = cp.async.bulk.tensor, SYNCS = mbarrier,")
    print("# HMMA = mma.sync, LDSM = ldmatrix.\n")
    name, counts, examples = None, None, None

    def flush():
        if name and any(counts[k] for k, _ in PATS if k != "SYNCS"):
            print(name)
            print("    " + "  ".join("%s=%d" % (k, counts[k]) for k, _ in PATS if counts[k]))
            for k, _ in PATS:
                if k in examples and k not in ("SYNCS",):
                    print("      e.g. " + examples[k])
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, counts, examples = m.group(1), collections.Counter(), {}
            continue
        if name is None:
            continue
        for k, pat in PATS:
            if re.search(pat, line):
                if k == "UTCMMA" and re.search(r"\bUTC[HQ]MMA\b", line):
                    continue
                counts[k] += 1
                examples.setdefault(k, " ".join(line.split())[:150])
    flush()


if __name__ == "__main__":
    sys.exit(main())
