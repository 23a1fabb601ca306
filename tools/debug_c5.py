import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2dsfs-scan_b200")); sys.path.insert(0, ROOT)
import torch, tdsfs_capi as T
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
cfg = dict(bench.WORKLOADS["config5"]); cfg["S"] = S
n1 = n2 = 500; w1 = w2 = 32; RW = 64
pos = np.concatenate(bench.positions_for(cfg, range(cfg["C"])))
off = np.concatenate([[0], np.cumsum(bench.chrom_sizes(S, cfg["C"]))]).astype(np.int64)
h = T.Handle(0); h.set_panel(n1, n2, True)
gw = (S + 31) // 32 * RW * 32
G = torch.empty((gw,), dtype=torch.int32, device="cuda")
h.synth_genotypes(G.data_ptr(), S, 0, w1, w2, n1, n2, 1); torch.cuda.synchronize(); print("synth ok", flush=True)
pd = torch.from_numpy(pos).cuda()
h.load_genotypes(G, S, w1, w2, n1, n2, pd, off); print("load ok", flush=True)
for it in range(3):
    h.background(T.BG_GENOME); print("k1 ok", h.timings()["k1_count"], flush=True)
    h.finalize_background(); print("fin ok", flush=True)
    n = h.scan(20000, fetch=False); print("scan ok", n, h.timings(), flush=True)
if S <= 400000:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sfs_oracle as O
    Gh = G.cpu().numpy().view(np.uint32)
    cnt = O.unpack_counts(Gh, w1, w2, n1, n2, S)
    e2, e1, e1b = O.dense_spectra(cnt, n1, n2)
    s2, s1a, s1b = h.get_background(0)
    print("bg equal:", np.array_equal(s2.astype(np.int64), e2), np.array_equal(s1a.astype(np.int64), e1), np.array_equal(s1b.astype(np.int64), e1b), int(e2.sum()), int(s2.sum()))
