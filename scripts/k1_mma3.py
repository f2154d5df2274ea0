import json, sys
d = json.load(open(sys.argv[1]))
mma = d["mma"]
chunks = [int(x) for x in sys.argv[2].split(",")]
per_tile = sum(3 * c + 1 for c in chunks)
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 0
i = tile * per_tile
base = mma[i]
for g, c in enumerate(chunks):
    rows = []
    for j in range(c):
        a, b, cc = mma[i], mma[i + 1], mma[i + 2]; i += 3
        rows.append((a - base, b - a, cc - b))
    commit = mma[i]; i += 1
    if g in (1, 2, 8, 9, 10):
        print(f"step {g}: (loop_top, act_wait, full_wait) " + " ".join(f"{t}/{w1}/{w2}" for t, w1, w2 in rows) + f" commit {commit-base}")
