import json, sys
d = json.load(open(sys.argv[1]))
tma, land, mma, ep = d["tma"], d["detail"], d["mma"], d["epilogue"]
chunks = [int(x) for x in sys.argv[2].split(",")]
per_tile = sum(chunks)
nt = min(len(tma), len(land)) // per_tile
off = (nt - 1) * per_tile + chunks[0] + chunks[1]      # step 2 of the last tile
ms = []; i = 0
for t in range(len(mma) // (per_tile + len(chunks))):
    for c in chunks:
        ms += mma[i:i + c]; i += c + 1
base = tma[off]
print("blk  issue  landed  lat | mma_ready  (cycles rel. to first issue of the step)")
for b in range(off, off + chunks[2] + 4):
    print(f"{b-off:3d} {tma[b]-base:7d} {land[b]-base:7d} {land[b]-tma[b]:6d} | {ms[b]-base:7d}")
