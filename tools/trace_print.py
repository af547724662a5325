#!/usr/bin/env python
"""print a device-side timeline written by bench.py --trace (one rank): microseconds since the step's first mark"""
import json, sys
NAMES = {1: "step start", 2: "buckets done", 3: "pass1 done", 4: "P exchange done", 5: "pass2 solids' share done", 6: "pass2 done",
         7: "sub-steps begin (side)", 8: "sub-steps end (side)", 9: "pre-step done (joined sub-steps)", 10: "migration+halo done"}
TAGS = ["vote", "mig", "halo", "P", "solP", "solV", "sub"]
PUSH = ["migL", "migR", "haloL", "haloR", "PL", "PR", "solP", "solV"]
def name(c):
    if c < 100: return NAMES.get(c, str(c))
    if c < 200: return "wait " + TAGS[c - 100] + " begins"
    if c < 300: return "wait " + TAGS[c - 200] + " ends"
    if c < 310: return "sub-step pass %d posted" % (c - 300 + 1)
    if c < 320: return "sub-step pass %d kernel starts" % (c - 310 + 1)
    if c < 400: return "sub-step pass %d last block done" % (c - 320 + 1)
    return "push " + PUSH[c - 400] + " complete"
d = json.load(open(sys.argv[1]))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 1
marks = sorted(d["marks"], key=lambda m: m[1])
starts = [t for c, t in marks if c == 1]
t0 = starts[step]; t1 = starts[step + 1] if step + 1 < len(starts) else marks[-1][1] + 1
print(f"rank {d['rank']} step {step}: {(t1 - t0) / 1e3:.1f} us")
prev = t0
for c, t in marks:
    if t0 <= t < t1 + 400000 and (t < t1 or c in (8,) or 300 <= c < 400 or c in (106, 206)):
        print(f"  {(t - t0) / 1e3:9.1f}  (+{(t - prev) / 1e3:7.1f})  {name(c)}")
        prev = t
