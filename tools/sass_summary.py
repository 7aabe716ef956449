"""Writes profiles/r2_sass_summary.txt: opcode histogram of the built library (cuobjdump -sass)."""
# see the inline generator in the session log; kept minimal: run
#   cuobjdump -sass gtn_applications_b200/lib/libwfst_b200.so | python tools/sass_summary.py
import collections, re, sys
txt = sys.stdin.read()
tot = collections.Counter()
for m in re.finditer(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', txt, re.M):
    tot[m.group(1).split('.')[0]] += 1
for k, v in tot.most_common(60):
    print("%-16s %d" % (k, v))
