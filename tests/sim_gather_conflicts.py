"""Design aid, not a test: shared-memory wavefront count of the streaming gather's LDS.128 under the lane -> voxel map
and the pixel-box row pitch, on the reference's exact indices (via the CPU oracle) for the Example rig.

Cost model = what tools/lds_bench.cu measures on the B200 (profiles/r02_lds_bench.txt): per quarter-warp one wavefront
per distinct 16-byte unit colliding in a bank group; two conflict-free quarters of a half-warp whose units do not collide
share one wavefront; identical addresses are free.

    python tests/sim_gather_conflicts.py zline 3        # first form of the kernel (8 voxels along z per quarter): 5.8 per load
    python tests/sim_gather_conflicts.py block224 3     # 2x2x2 cube per quarter, pitch == 3 (mod 8): 3.3 per load
    python tests/sim_gather_conflicts.py best 2,3,5,6   # best residue per (tile, camera): 2.9 per load
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _indices():
    import jarvis_hybridnet_b200.synth as S
    from oracle import hybridnet_oracle as O
    sh = S.EXAMPLE
    cam, intr, dist = S.make_rig(sh.ncam, 0)
    hm, c3, chm, _ = S.make_frameset(sh, cam, intr, dist, 0)
    idx = O.reproject_indices(c3, chm, cam, intr, dist, sh.G, sh.spacing, sh.hs)
    return (idx % sh.hs).astype(np.int64), (idx // sh.hs).astype(np.int64)


PX, PY = _indices()
ncam, G = PX.shape[0], PX.shape[1]

def qcost(p, act):
    """p, act: [N, 8] -> per-quarter cost [N] and 'clean' flag (<=1 distinct per group)"""
    N = p.shape[0]
    pm = (p >> 20) & 7
    cost = np.zeros(N, np.int64)
    for b in range(8):
        sel = (pm == b) & act
        cnt = np.zeros(N, np.int64)
        for i in range(8):
            new = sel[:, i].copy()
            for j in range(i):
                new &= ~(sel[:, j] & (p[:, j] == p[:, i]))
            cnt += new
        cost = np.maximum(cost, cnt)
    return cost

def hcost(p, act):
    """p, act: [N,16] half-warps -> wavefronts [N] under the merge model"""
    c0, c1 = qcost(p[:, :8], act[:, :8]), qcost(p[:, 8:], act[:, 8:])
    # union clean?
    N = p.shape[0]
    pm = (p >> 20) & 7
    union = np.zeros(N, np.int64)
    for b in range(8):
        sel = (pm == b) & act
        cnt = np.zeros(N, np.int64)
        for i in range(16):
            new = sel[:, i].copy()
            for j in range(i):
                new &= ~(sel[:, j] & (p[:, j] == p[:, i]))
            cnt += new
        union = np.maximum(union, cnt)
    merged = (union <= 1)
    return np.where(merged, np.minimum(union, 1), c0 + c1), merged

def tile_arrays(c, ti, tj, tk, TZ=8):
    I = 8 * ti - 1 + np.arange(8); J = 8 * tj - 1 + np.arange(8); K = TZ * tk + np.arange(TZ)
    vi = (I >= 0) & (I < G); vj = (J >= 0) & (J < G)
    Ic, Jc = np.clip(I, 0, G - 1), np.clip(J, 0, G - 1)
    x = PX[c][np.ix_(Ic, Jc, K)]; y = PY[c][np.ix_(Ic, Jc, K)]
    act = np.broadcast_to(vi[:, None, None] & vj[None, :, None], x.shape)
    return x, y, act

def pick_r(x, y, act, mode):
    if isinstance(mode, int): return mode
    if mode == 'extent':      # pitch residue = x-extent of a 2x2x2 cube footprint + 1 (so rows of the footprint do not collide)
        w = 1
        xs = x.reshape(4, 2, 4, 2, -1, 2)   # li=(cx,2), lj=(cy,2), lk=(cz,2)
        ext = (xs.max(axis=(1, 3, 5)) - xs.min(axis=(1, 3, 5))).max()
        return int(ext) + 1
    raise ValueError

def eval_best(mapping, cands):
    tot = 0; nhalf = 0
    nt = G // 8 + 1
    hist = {}
    for c in range(ncam):
        for ti in range(nt):
            for tj in range(nt):
                for tk in range(G // 8):
                    x, y, act = tile_arrays(c, ti, tj, tk)
                    best = None
                    for r in cands:
                        p = (((x + r * y) & 7) << 20) | (y << 10) | x
                        pb = p.reshape(4, 2, 4, 2, 2, 4).transpose(0, 2, 4, 5, 3, 1)
                        ab = act.reshape(4, 2, 4, 2, 2, 4).transpose(0, 2, 4, 5, 3, 1)
                        P_ = pb.reshape(-1, 16); A_ = ab.reshape(-1, 16)
                        w, m = hcost(P_, A_)
                        anyact = A_.any(1)
                        t = int(w[anyact].sum())
                        if best is None or t < best[0]: best = (t, int(anyact.sum()), r)
                    tot += best[0]; nhalf += best[1]; hist[best[2]] = hist.get(best[2], 0) + 1
    return tot, nhalf, hist

def eval_mapping(mapping, mode):
    tot = 0; ninstr = 0; nmerged = 0; nhalf = 0
    nt = G // 8 + 1
    for c in range(ncam):
        for ti in range(nt):
            for tj in range(nt):
                for tk in range(G // 8):
                    x, y, act = tile_arrays(c, ti, tj, tk)
                    r = pick_r(x, y, act, mode)
                    p = (((x + r * y) & 7) << 20) | (y << 10) | x      # bank residue in the high bits, pixel identity below
                    if mapping == 'zline':      # half-warp = 2 y-pairs? current: lane = cj*8 + z, warp = ci; voxel v=(iv,jv): half-warp lanes 0-15: cj in {0,1}
                        # instruction (ci, v): lanes (cj, z) -> voxel (2ci+iv, 2cj+jv, z)
                        halves = []
                        for ci in range(4):
                            for iv in range(2):
                                for jv in range(2):
                                    for h2 in range(2):
                                        lj = [2 * (2 * h2) + jv, 2 * (2 * h2 + 1) + jv]
                                        pp = np.concatenate([p[2 * ci + iv, lj[0]], p[2 * ci + iv, lj[1]]])
                                        aa = np.concatenate([act[2 * ci + iv, lj[0]], act[2 * ci + iv, lj[1]]])
                                        halves.append((pp, aa))
                        P_ = np.stack([h[0] for h in halves]); A_ = np.stack([h[1] for h in halves])
                    elif mapping == 'block224':  # half-warp = 2x2x4 voxel block: lane = (bz*2+by)*2+bx ; quarter = cube
                        pb = p.reshape(4, 2, 4, 2, 2, 4).transpose(0, 2, 4, 5, 3, 1)   # [cx, cy, zb, bz(4), by(2), bx(2)]
                        ab = act.reshape(4, 2, 4, 2, 2, 4).transpose(0, 2, 4, 5, 3, 1)
                        P_ = pb.reshape(-1, 16); A_ = ab.reshape(-1, 16)
                    elif mapping == 'stride2':   # thread = 2x2 (x,y) at one z (current); half-warp = {ci,ci+1} x {cj,cj+1} x 4 z
                        halves = []
                        for iv in range(2):
                            for jv in range(2):
                                for cib in range(2):
                                    for cjb in range(2):
                                        for zb in range(2):
                                            xs = [2 * (2 * cib + a) + iv for a in range(2)]
                                            ys = [2 * (2 * cjb + b) + jv for b in range(2)]
                                            zs = list(range(4 * zb, 4 * zb + 4))
                                            pp = p[np.ix_(xs, ys, zs)].reshape(-1); aa = act[np.ix_(xs, ys, zs)].reshape(-1)
                                            halves.append((pp, aa))
                        P_ = np.stack([h[0] for h in halves]); A_ = np.stack([h[1] for h in halves])
                    elif mapping == 'block422':  # half-warp = 4(x) x 2 x 2
                        pb = p.reshape(2, 4, 4, 2, 4, 2).transpose(0, 2, 4, 5, 3, 1)   # [xb, cy, cz, bz(2), by(2), bx(4)]
                        ab = act.reshape(2, 4, 4, 2, 4, 2).transpose(0, 2, 4, 5, 3, 1)
                        P_ = pb.reshape(-1, 16); A_ = ab.reshape(-1, 16)
                    w, m = hcost(P_, A_)
                    anyact = A_.any(1)
                    tot += int(w[anyact].sum()); nhalf += int(anyact.sum()); nmerged += int((m & anyact).sum())
    return tot, nhalf, nmerged

if __name__ == '__main__' and sys.argv[1] == 'best':
    tot, nhalf, hist = eval_best('block224', [int(v) for v in sys.argv[2].split(',')])
    print('best of', sys.argv[2], 2 * tot / nhalf, hist)
    sys.exit()
if __name__ == '__main__':
    mapping, mode = sys.argv[1], sys.argv[2]
    mode = int(mode) if mode.isdigit() else mode
    tot, nhalf, nm = eval_mapping(mapping, mode)
    print(mapping, mode, 'wavefronts per half-warp-instr', tot / nhalf, 'merged share', nm / nhalf, '-> per LDS.128 (2 halves):', 2 * tot / nhalf)
