"""TEST INFRASTRUCTURE: a NumPy emulation, on HOST pointers, of the libregat.so entry points that
tf_vqa_regat_b200/question.py sequences (regat_gemm with its alpha / bias / accumulate epilogue and the regat_q_* kernels,
semantics as documented in include/regat.h).  It lets the CPU suite dry-run the module's launch sequence -- operand
shapes, transposes, leading dimensions, buffer offsets, the order of the back-propagation through time -- against the
oracle without a GPU.  It is not a fallback: the product never imports it, and the CUDA kernels are checked against the
oracle on the GPU by tests/test_gpu_question.py."""
import ctypes as C

import numpy as np


def _arr(ptr, n, dtype=np.float32):
    ct = {np.float32: C.c_float, np.int32: C.c_int32}[dtype]
    return np.ctypeslib.as_array((ct * int(n)).from_address(int(ptr)))


def _mat(ptr, rows, cols, ld):
    if rows == 0 or cols == 0:
        return np.zeros((rows, cols), np.float32)
    base = _arr(ptr, (rows - 1) * ld + cols)
    return np.lib.stride_tricks.as_strided(base, shape=(rows, cols), strides=(4 * ld, 4))


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


class HostOps:
    def __init__(self):
        self.calls = []

    def regat_gemm(self, dtype, tA, tB, M, N, K, A, lda, B, ldb, Cp, ldc, c_dtype, epi, stream):
        assert dtype == 0 and c_dtype == 0
        assert lda >= (M if tA else K) and ldb >= (K if tB else N) and ldc >= N, "leading dimension too small"
        a = _mat(A, K, M, lda).T if tA else _mat(A, M, K, lda)
        b = _mat(B, N, K, ldb).T if tB else _mat(B, K, N, ldb)
        x = a.astype(np.float64) @ b.astype(np.float64)
        e = epi._obj if epi is not None else None
        c = _mat(Cp, M, N, ldc)
        if e is not None:
            assert not e.addend and not e.relu and not e.gate and not e.c2 and e.split_k <= 1
            if e.alpha:
                assert e.alpha_cols == 0
                x = x * float(_arr(e.alpha, 1)[0])
            if e.bias:
                x = x + _arr(e.bias, N).astype(np.float64)[None, :]
            if e.accumulate:
                x = x + c.astype(np.float64)
        c[...] = x.astype(np.float32)
        self.calls.append("gemm")
        return 0

    def regat_q_embed_fwd(self, tokens, BT, n_token, E, emb, emb2, out, stream):
        tok = _arr(tokens, BT, np.int32)
        W = 2 * E if emb2 else E
        o = _mat(out, BT, W, W)
        t1 = _mat(emb, n_token + 1, E, E)
        o[:, :E] = t1[tok] * (tok != n_token)[:, None]
        if emb2:
            o[:, E:] = _mat(emb2, n_token + 1, E, E)[tok] * (tok != n_token)[:, None]
        return 0

    def regat_q_embed_bwd(self, tokens, BT, n_token, E, width, dX, demb, demb2, stream):
        tok = _arr(tokens, BT, np.int32)
        d = _mat(dX, BT, width, width)
        keep = tok != n_token
        if demb:
            np.add.at(_mat(demb, n_token + 1, E, E), tok[keep], d[keep, :E])
        if demb2:
            np.add.at(_mat(demb2, n_token + 1, E, E), tok[keep], d[keep, E:])
        return 0

    def regat_q_gru_gates_fwd(self, B, H, xi, ld_xi, hi, hp, ld_hp, h_out, ld_h, z, r, c, hp_copy, stream):
        x, h3 = _mat(xi, B, 3 * H, ld_xi).astype(np.float64), _mat(hi, B, 3 * H, 3 * H).astype(np.float64)
        hprev = _mat(hp, B, H, ld_hp).astype(np.float64) if hp else np.zeros((B, H))
        zz = _sig(x[:, :H] + h3[:, :H]); rr = _sig(x[:, H:2 * H] + h3[:, H:2 * H])
        cc = np.tanh(x[:, 2 * H:] + rr * h3[:, 2 * H:])
        _mat(h_out, B, H, ld_h)[...] = zz * hprev + (1 - zz) * cc
        _mat(z, B, H, H)[...] = zz; _mat(r, B, H, H)[...] = rr; _mat(c, B, H, H)[...] = cc
        if hp_copy:
            _mat(hp_copy, B, H, H)[...] = hprev
        return 0

    def regat_q_gru_gates_bwd(self, B, H, dh_seq, ld_dseq, dh_rec, z, r, c, hp_copy, hi, dxi, ld_dxi, dhi, dhp, stream):
        dh = _mat(dh_seq, B, H, ld_dseq).astype(np.float64) + (_mat(dh_rec, B, H, H).astype(np.float64) if dh_rec else 0.0)
        zz, rr, cc, hp = (_mat(p, B, H, H).astype(np.float64) for p in (z, r, c, hp_copy))
        hh = _mat(hi, B, 3 * H, 3 * H).astype(np.float64)[:, 2 * H:]
        da = dh * (1 - zz) * (1 - cc * cc)
        dzp = dh * (hp - cc) * zz * (1 - zz)
        drp = da * hh * rr * (1 - rr)
        dx, d3 = _mat(dxi, B, 3 * H, ld_dxi), _mat(dhi, B, 3 * H, 3 * H)
        dx[:, :H] = dzp; dx[:, H:2 * H] = drp; dx[:, 2 * H:] = da
        d3[:, :H] = dzp; d3[:, H:2 * H] = drp; d3[:, 2 * H:] = da * rr
        _mat(dhp, B, H, H)[...] = dh * zz
        return 0

    def regat_q_tanh_fwd(self, x, n, stream):
        a = _arr(x, n); a[...] = np.tanh(a.astype(np.float64))
        return 0

    def regat_q_tanh_bwd(self, dy, y, n, stream):
        d, yy = _arr(dy, n), _arr(y, n).astype(np.float64)
        d[...] = d * (1 - yy * yy)
        return 0

    def regat_q_batch_softmax_fwd(self, logits, B, T, P, stream):
        lt = _mat(logits, B, T, T).astype(np.float64).T                    # [T,B]
        e = np.exp(lt - lt.max(1, keepdims=True))
        _mat(P, T, B, B)[...] = e / e.sum(1, keepdims=True)
        return 0

    def regat_q_batch_softmax_bwd(self, P, dP, B, T, dlogits, stream):
        p, d = _mat(P, T, B, B).astype(np.float64), _mat(dP, T, B, B).astype(np.float64)
        _mat(dlogits, B, T, T)[...] = (p * (d - (p * d).sum(1, keepdims=True))).T
        return 0

    def regat_q_pool_fwd(self, w_flat, seq, B, T, H, q_att, stream):
        w, s = _mat(w_flat, B, T, T).astype(np.float64), _arr(seq, B * T * H).reshape(B, T, H).astype(np.float64)
        _mat(q_att, B, H, H)[...] = np.einsum("bt,bth->bh", w, s)
        return 0

    def regat_q_pool_bwd(self, w_flat, seq, dq_att, dq_last, B, T, H, dseq, dW, stream):
        w, s = _mat(w_flat, B, T, T).astype(np.float64), _arr(seq, B * T * H).reshape(B, T, H).astype(np.float64)
        g = _mat(dq_att, B, H, H).astype(np.float64)
        d = w[:, :, None] * g[:, None, :]
        if dq_last:
            d[:, T - 1] += _mat(dq_last, B, H, H)
        _arr(dseq, B * T * H)[...] = d.reshape(-1)
        _mat(dW, B, T, T)[...] = np.einsum("bh,bth->bt", g, s)
        return 0

    def regat_q_dot(self, a, b, n, out, stream):
        _arr(out, 1)[0] += np.float32(_arr(a, n).astype(np.float64) @ _arr(b, n).astype(np.float64))
        return 0

    def regat_q_wn_alpha(self, g, vv, alpha, stream):
        _arr(alpha, 1)[0] = _arr(g, 1)[0] / np.sqrt(max(float(_arr(vv, 1)[0]), 1e-12))
        return 0

    def regat_q_wn_bwd(self, G, v, g, vv, Gv, n, dv, dg, stream):
        nn = max(float(_arr(vv, 1)[0]), 1e-12)
        alpha, k = float(_arr(g, 1)[0]) / np.sqrt(nn), float(_arr(Gv, 1)[0]) / nn
        _arr(dv, n)[...] = alpha * (_arr(G, n).astype(np.float64) - k * _arr(v, n).astype(np.float64))
        _arr(dg, 1)[0] = float(_arr(Gv, 1)[0]) / np.sqrt(nn)
        return 0

    def regat_q_clip_adamax(self, w, grad, m, u, n, gsumsq, clip, lr, step, b1, b2, eps, stream):
        g = _arr(grad, n).astype(np.float64) * (clip / max(np.sqrt(float(_arr(gsumsq, 1)[0])), clip))
        mm, uu, ww = _arr(m, n), _arr(u, n), _arr(w, n)
        mi = mm + (g - mm) * (1 - b1)
        ui = np.maximum(b2 * uu.astype(np.float64), np.abs(g))
        mm[...] = mi; uu[...] = ui
        ww[...] = ww - (lr / (1 - b1 ** step)) * mi / (ui + eps)
        return 0

    def regat_q_embed_sumsq(self, tokens, BT, n_token, E, width, col0, dX, out, stream):
        tok = _arr(tokens, BT, np.int32)
        d = _mat(dX, BT, width, width)[:, col0:col0 + E].astype(np.float64)
        _arr(out, 1)[0] += float((d[tok != n_token] ** 2).sum())
        return 0

    def regat_q_embed_clip_adamax(self, tokens, BT, n_token, E, width, col0, dX, table, grad, m, u, uinc, gsumsq, clip, lr, step, b1, b2, eps,
                                  stream):
        tok = _arr(tokens, BT, np.int32)
        n = (n_token + 1) * E
        scale = clip / max(np.sqrt(float(_arr(gsumsq, 1)[0])), clip)
        vals = _mat(dX, BT, width, width)[:, col0:col0 + E].astype(np.float64) * scale
        uu = _arr(u, n).reshape(n_token + 1, E)
        ud = b2 * uu.astype(np.float64)
        inc = np.zeros_like(ud)
        live = tok != n_token
        np.add.at(inc, tok[live], np.maximum(ud[tok[live]], np.abs(vals[live])) - ud[tok[live]])
        g = _arr(grad, n).astype(np.float64) * scale
        mm, ww = _arr(m, n), _arr(table, n)
        mi = mm + (g - mm) * (1 - b1)
        ui = (ud + inc).ravel()
        mm[...] = mi; uu[...] = ui.reshape(uu.shape)
        ww[...] = ww - (lr / (1 - b1 ** step)) * mi / (ui + eps)
        _arr(uinc, n)[...] = 0
        return 0

