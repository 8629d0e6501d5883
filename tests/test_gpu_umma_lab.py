"""GPU (-m gpu): every tcgen05 operand flavour the tensor-core engine uses, isolated in the UMMA lab and checked
against numpy on fp16-rounded operands."""
import numpy as np
import pytest

from _umma import CC, f16, idesc, put_chunkcols, put_kmajor, run_lab

pytestmark = pytest.mark.gpu
TOL = 2e-3   # fp32 accumulation of <= 128 products of O(1) fp16 values


def test_kmajor_projection_n96():
    rng = np.random.default_rng(0)
    a, w = rng.standard_normal((128, 96)), rng.standard_normal((96, 96)) * 0.3       # w: [N][K]
    img = np.zeros(64 * 1024, dtype=np.uint8)
    put_chunkcols(img, 0, a)
    put_kmajor(img, 32768, w, 12 * 128, 128)
    ops = [(ks * 2 * CC, CC, 128, 32768 + ks * 2 * 1536, 1536, 128, idesc(96), 0, int(ks > 0)) for ks in range(6)]
    d = run_lab(img, ops, 96)
    assert np.abs(d - f16(a) @ f16(w).T).max() < TOL


@pytest.mark.parametrize("h", [0, 1, 3])
def test_scores_with_zero_chunk(h):
    """S_h = Q_h K_h^T for a 24-wide head: second K=16 step pairs chunk 3h+2 with a shared zero chunk through a
    per-descriptor leading-dimension offset."""
    rng = np.random.default_rng(1)
    q, k = rng.standard_normal((128, 96)), rng.standard_normal((128, 96))
    img = np.zeros(96 * 1024, dtype=np.uint8)
    QO, KO = 0, 13 * CC                      # Q: 12 chunk columns + 1 zero chunk column; K (+V-like finite data after it)
    put_chunkcols(img, QO, q)
    put_chunkcols(img, KO, np.concatenate([k, rng.standard_normal((128, 8))], axis=1))   # finite garbage after K
    ops = [(QO + 3 * h * CC, CC, 128, KO + 3 * h * CC, CC, 128, idesc(128), 0, 0),
           (QO + (3 * h + 2) * CC, (12 - 3 * h - 2) * CC, 128, KO + (3 * h + 2) * CC, CC, 128, idesc(128), 0, 1)]
    d = run_lab(img, ops, 128)
    ref = f16(q[:, 24 * h:24 * h + 24]) @ f16(k[:, 24 * h:24 * h + 24]).T
    assert np.abs(d - ref).max() < TOL


@pytest.mark.parametrize("h", [0, 2])
def test_pv_mn_major_b(h):
    """O_h = P V_h with V in the activation layout consumed as an MN-major B operand (N = 32 channels from 24h)."""
    rng = np.random.default_rng(2)
    p, v = rng.random((128, 128)), rng.standard_normal((128, 104))
    img = np.zeros(96 * 1024, dtype=np.uint8)
    PO, VO = 0, 16 * CC
    put_chunkcols(img, PO, p)
    put_chunkcols(img, VO, v)
    # MN-major B: stride-dimension offset = distance between 8-channel groups (CC), leading-dimension = between 8-row groups (128)
    ops = [(PO + s * 2 * CC, CC, 128, VO + 3 * h * CC + s * 256, 128, CC, idesc(32, b_mn=True), 0, int(s > 0)) for s in range(8)]
    d = run_lab(img, ops, 32)
    ref = f16(p) @ f16(v[:, 24 * h:24 * h + 32])
    assert np.abs(d - ref).max() < TOL


@pytest.mark.parametrize("pose", [0, 3, 6])
def test_transposed_pose_aggregation_mn_major_a(pose):
    """D^T[c][i] = sum_j Y[17p+j][c] L[i][j]: activations read as an MN-major A operand starting at an arbitrary row,
    the 17x17 matrix zero-padded to a 32x32 K-major B operand."""
    rng = np.random.default_rng(3)
    y = rng.standard_normal((128, 128))       # 96 real channels + what lies behind them in memory (finite)
    lmat = rng.standard_normal((17, 17))
    lpad = np.zeros((32, 32)); lpad[:17, :17] = lmat
    img = np.zeros(96 * 1024, dtype=np.uint8)
    YO, LO = 0, 40 * 1024
    put_chunkcols(img, YO, y)
    # rows 128..143 of the first chunk columns alias the start of the next ones (16-B skew): keep them finite (they are)
    put_kmajor(img, LO, lpad, 512, 128)
    ops = [(YO + (17 * pose + 16 * s) * 16, 128, CC, LO + s * 2 * 512, 512, 128, idesc(32, a_mn=True), 0, int(s > 0)) for s in range(2)]
    d = run_lab(img, ops, 32)
    ref = f16(y[17 * pose:17 * pose + 17, :]).T @ f16(lmat).T        # [128 ch][17]
    assert np.abs(d[:96, :17] - ref[:96]).max() < TOL


def _put_tall(img, off, g, lbo=4112):
    """G [17][17] -> rows 128..144 of a tall K-major operand [256 rows x 32], element (r, k) at off + (k//8)*lbo + r*16 + (k%8)*2"""
    v = img.view(np.float16)
    for i in range(17):
        for k in range(17):
            v[(off + (k // 8) * lbo + (128 + i) * 16 + (k % 8) * 2) // 2] = np.float16(g[i, k])


@pytest.mark.parametrize("poses", [[0], [3], [6], list(range(7))])
def test_tall_window_graph_aggregation(poses):
    """OUT = sum_p window_p(G) * ACT[17p .. 17p+31]: the K-major A operand is a 128-row window (16-byte granular start)
    into a zero-padded tall operand, the activations are an MN-major B operand starting at row 17p (dp_tc2.cu)."""
    rng = np.random.default_rng(4)
    g = rng.standard_normal((17, 17))
    act = rng.standard_normal((128, 96))
    img = np.zeros(96 * 1024, dtype=np.uint8)
    AO, TO, TL = 0, 40 * 1024, 4112
    put_chunkcols(img, AO, act)
    _put_tall(img, TO, g, TL)
    ops = []
    for p in poses:
        for s in range(2):
            ops.append((TO + (128 - 17 * p) * 16 + s * 2 * TL, TL, 128, AO + (17 * p + 16 * s) * 16, 128, CC, idesc(96, b_mn=True), 0, int(len(ops) > 0)))
    d = run_lab(img, ops, 96)
    ref = np.zeros((128, 96))
    for p in poses:
        ref[17 * p:17 * p + 17] = f16(g) @ f16(act[17 * p:17 * p + 17])
    assert np.abs(d - ref).max() < TOL


@pytest.mark.parametrize("h", [0, 3])
def test_pv_with_p_in_tmem(h):
    """O_h = P V_h with P (fp16, two per 32-bit column, lane = row) read from tensor memory as the A operand and V in
    the activation layout consumed as an MN-major B operand."""
    from _umma import run_lab_ts
    rng = np.random.default_rng(5)
    p, v = rng.random((128, 128)), rng.standard_normal((128, 104))
    img = np.zeros(64 * 1024, dtype=np.uint8)
    put_chunkcols(img, 0, v)
    timg = p.astype(np.float16).view(np.uint32).reshape(128, 64)     # little endian: low half = even k
    TC0 = 256
    ops = [(TC0 + 8 * s, 0, 0, 3 * h * CC + s * 256, 128, CC, idesc(32, b_mn=True), 0, 2 | int(s > 0)) for s in range(8)]
    d = run_lab_ts(img, timg, TC0, ops, 32)
    ref = f16(p) @ f16(v[:, 24 * h:24 * h + 32])
    assert np.abs(d - ref).max() < TOL
