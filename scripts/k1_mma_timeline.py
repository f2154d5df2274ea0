"""Print the MMA-thread view of gpurun_out/k1_timeline_*.json: one stamp per K chunk (after the full-barrier wait)
and one per step (after the accumulator commit)."""
import json, sys
d = json.load(open(sys.argv[1]))
mma, ep = d["mma"], d["epilogue"]
chunks = [int(x) for x in sys.argv[2].split(",")]      # chunks per step (parts * k-chunks)
per_tile = sum(c + 1 for c in chunks)
nt = len(mma) // per_tile
k = (nt - 1) * per_tile
pe = 2 + 2 * len(chunks)
ke = (len(ep) // pe - 1) * pe
base = ep[ke]
i = k
for g, c in enumerate(chunks):
    ts = [mma[i + j] - base for j in range(c)]
    commit = mma[i + c] - base
    i += c + 1
    ea, eb = ep[ke + 2 + 2 * g] - base, ep[ke + 3 + 2 * g] - base
    d_ = [ts[j + 1] - ts[j] for j in range(len(ts) - 1)]
    print(f"step {g:2d}: epi [{ea:7d},{eb:7d}] first chunk {ts[0]:7d} deltas {d_} commit {commit}")
