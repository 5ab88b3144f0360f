"""The certificate of the tensor-core input layer (csrc/input_tc.cu), restated in numpy and checked against the oracle:
whenever the error bound says "no half-integer within D of z", the oracle's layer-0 byte (the reference's exact fp32
arithmetic, src/cpp/dnn.cc:175-286) must be the bucket the certificate predicts — and the bound must leave only a few
per cent of the elements undecided, or the fast path would not be one.  No GPU needed: this pins the mathematics."""
import numpy as np
import pytest

from fast_dnn_b200 import synth
import oracle_py

U = 2.0 ** -24


def round_counts(I):
    """roundings term k goes through in the reference: product, later adds of its SSE lane, two combining adds (+1)"""
    n = I // 4
    q = np.arange(I) // 4
    return np.where(q == 0, n - 1, n - q) + 3 + 1


def certificate(x_raw, w, bias, shift, scale):
    """→ (k, certain) per (frame, node), following input_tc.cu's derivation with exact integers and fp64"""
    xq = ((x_raw.astype(np.float32) + shift.astype(np.float32)).astype(np.float32) * scale.astype(np.float32)).astype(np.float32)
    I = xq.shape[1]
    c = round_counts(I).astype(np.float64)

    def fixed(v):
        mx = np.abs(v).max(axis=1)
        e = np.where(mx > 0, np.floor(np.log2(np.maximum(mx, 1e-300))) + 1, 0).astype(np.int64)
        e = np.where(2.0 ** e <= mx, e + 1, e)  # mx < 2^e
        sc = 2.0 ** (e - 22)
        V = np.rint(v.astype(np.float64) / sc[:, None]).astype(np.int64)
        assert np.abs(V).max() <= 2 ** 22
        return V, sc

    X, sx = fixed(xq)
    W, sw = fixed(w)
    w_low = ((W + 128) & 255) - 128  # the weights' lowest BALANCED digit (an s8; fdnn_api.cu: upload_model)
    low = (X & 255) @ w_low.T  # the product of the lowest limbs, which the kernel bounds instead of computing
    total = X @ W.T - low
    z = (total * np.outer(sx, sw) + bias.astype(np.float64)[None, :]) * 100.0
    nxc = np.sqrt((c * xq.astype(np.float64) ** 2).sum(axis=1))
    nwc = np.sqrt((c * w.astype(np.float64) ** 2).sum(axis=1))
    nx = np.sqrt((xq.astype(np.float64) ** 2).sum(axis=1)) + sx * 131072.0 * np.sqrt(I)
    nw = np.sqrt((w.astype(np.float64) ** 2).sum(axis=1)) + sw * 131072.0 * np.sqrt(I)
    eps_q = np.outer(sx, sw) * (0.5 * np.abs(X).sum(axis=1)[:, None] + 0.5 * np.abs(W).sum(axis=1)[None, :] + 0.25 * I
                                + 255.0 * np.abs(w_low).sum(axis=1)[None, :])
    bc = np.abs(bias.astype(np.float64) * 100.0)
    D = (100.0 * (U * 1.0001 * np.outer(nxc, nwc) + 7 * U * np.outer(nx, nw) + eps_q) + 3.1 * U * np.abs(z) + U * bc[None, :] + 2.1 * U + 1e-9) * 1.00001
    k = np.rint(z)
    certain = (np.abs(z - k) + D < 0.5) & (np.abs(k) < 640)
    certain |= (z - D > 639.5) | (z + D < -639.5)
    return np.clip(k, -640, 640).astype(np.int64), certain


@pytest.mark.parametrize("shape,n,stress", [("S", 96, False), ("P", 64, False), ("S", 64, True)])
def test_certified_buckets_equal_the_oracle(net_file, shape, n, stress):
    port = oracle_py.Port(net_file(shape, stress=stress))
    w, b, sh, sc = port.input_layer()
    frames = synth.make_frames(n, w.shape[1], seed=17)
    want = port.hidden_trace(frames)[0]  # u8 [n][H] after layer 0
    k, certain = certificate(frames, w, b, sh, sc)
    lut = oracle_py.Port.sigmoid_lut()
    predicted = np.where(k <= -640, 0, np.where(k >= 640, 255, lut[np.clip(k + 640, 0, 1279)]))
    wrong = certain & (predicted != want)
    assert not wrong.any(), f"{wrong.sum()} certified elements differ from the reference"
    assert certain.mean() > 0.9, f"only {100 * certain.mean():.1f} % decided: the bound is too loose to be useful"


def test_round_counts_match_the_reference_loop():
    # dnn.cc:219-247: lane r adds terms k = r, r+4, …; the first add of a lane is 0 + p (exact)
    c = round_counts(440)
    assert c[0] == c[3] == 109 + 4 and c[4] == 109 + 4 and c[8] == 108 + 4 and c[436] == c[439] == 1 + 4
