"""Per-step opcode counts of k_disk variants: python tools/sass_compare.py PERIOD a.sass [b.sass ...]
(b - a columns are printed when two files are given: e.g. steady copy = with_steady - generic_only)."""
import re, sys, collections
ALU = {'FMNMX3', 'FMNMX', 'ISETP', 'VIADD', 'VIMNMX3', 'IADD3', 'SHF', 'LOP3', 'LEA', 'FSEL', 'SEL', 'PLOP3', 'VIMNMX',
       'IABS', 'SGXT', 'VOTE', 'VOTEU', 'PRMT', 'FSETP', 'FCHK', 'BMSK', 'FLO', 'POPC'}


def hist(f):
    c = collections.Counter()
    for l in open(f):
        m = re.match(r'^\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P[0-9T]+ )?([A-Z0-9_]+)', l)
        if m:
            c[m.group(2)] += 1
    return c


if __name__ == "__main__":
    per = float(sys.argv[1])
    hs = [hist(f) for f in sys.argv[2:]]
    cols = hs + ([collections.Counter({k: hs[1][k] - hs[0][k] for k in hs[1]})] if len(hs) == 2 else [])
    tot = [0.0] * len(cols); alu = [0.0] * len(cols)
    for op in sorted(set().union(*cols), key=lambda k: -cols[-1][k]):
        v = [c[op] / per for c in cols]
        for i, x in enumerate(v):
            tot[i] += x
            if op in ALU:
                alu[i] += x
        if max(abs(x) for x in v) >= 0.5:
            print("%-12s" % op + "".join("%9.1f" % x for x in v))
    print("%-12s" % "total" + "".join("%9.1f" % x for x in tot))
    print("%-12s" % "alu pipe" + "".join("%9.1f" % x for x in alu))
