"""Host emulator of the C ABI -- TEST INFRASTRUCTURE ONLY.

Each method restates, with plain torch CPU ops, what the CUDA entry point of the same name in
include/dgmr_b200.h computes on the same arguments (in-place on the output tensors).  It serves two
purposes and is never reachable from the product package:
  * `-m "not gpu"` tests inject it with `_lib.set_backend()` to exercise the HOST logic (module wiring,
    autograd plumbing, state-dict contract, group/timestep folding) against the oracle without a GPU;
  * `-m gpu` kernel tests use it as the per-kernel reference for the CUDA kernels.
"""
import math

import torch
import torch.nn.functional as F


def _rows(t, rows, ld, ch):
    """[rows, ch] window with row pitch ld starting at t's first element (t may be a 1-D tail slice of a wider buffer)."""
    return torch.as_strided(t, (rows, ch), (ld, 1), t.storage_offset())


def _view(t, off, shape, strides):
    flat = t.reshape(-1)  # contiguous input: a view; as_strided offsets are absolute in the storage
    return torch.as_strided(flat, tuple(shape), tuple(strides), flat.storage_offset() + off)


class EmuBackend:
    name = "emu"

    def __init__(self):
        self.launches = 0

    # ---- queries
    def conv_umma_supported(self, *a):
        return False

    def wgrad_umma_supported(self, *a):
        return False

    # ---- layout
    def permute(self, src, dst, shape, sstr, dstr, accumulate=False, src_off=0, dst_off=0):
        s = _view(src, src_off, shape, sstr)
        d = _view(dst, dst_off, shape, dstr)
        if accumulate:
            d.add_(s)
        else:
            d.copy_(s)

    def reduce_mid(self, x, y, A, R, C, accumulate=False):
        s = x.reshape(A, R, C).sum(1)
        if accumulate:
            y.reshape(A, C).add_(s)
        else:
            y.reshape(A, C).copy_(s)

    # ---- pointwise
    def axpby(self, a, x, b, y, out):
        r = a * x if y is None else a * x + b * y
        out.copy_(r)

    def fill(self, x, value):
        x.fill_(value)

    def relu_fwd(self, x, y):
        y.copy_(torch.relu(x))

    def relu_bwd(self, dy, x, dx):
        dx.copy_(dy * (x > 0))

    @staticmethod
    def _rna_tf32(x):
        u = x.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF        # raw bit pattern (sign-magnitude)
        r = (u + 0x1000) & 0xFFFFE000                                              # nearest, ties away in magnitude (cvt.rna)
        r = torch.where(r >= 2 ** 31, r - 2 ** 32, r).to(torch.int32)
        return r.view(torch.float32)

    def round_tf32(self, x, y=None):
        (x if y is None else y).copy_(self._rna_tf32(x))

    def split_tf32(self, x, hi, lo):
        h = self._rna_tf32(x)
        hi.copy_(h)
        lo.copy_(self._rna_tf32(x - h))

    def pool_sum(self, x, y, N, D, H, W, C, pd, ph, pw, scale):
        v = x.reshape(N, D, H, W, C)[:, :D // pd * pd, :H // ph * ph, :W // pw * pw]
        v = v.reshape(N, D // pd, pd, H // ph, ph, W // pw, pw, C).sum(dim=(2, 4, 6)) * scale
        y.copy_(v.reshape(y.shape))

    def upsample(self, x, y, N, D, H, W, C, ud, uh, uw, Do, Ho, Wo, scale):
        v = x.reshape(N, D, H, W, C).repeat_interleave(ud, 1).repeat_interleave(uh, 2).repeat_interleave(uw, 3) * scale
        out = torch.zeros(N, Do, Ho, Wo, C, dtype=x.dtype)
        out[:, :D * ud, :H * uh, :W * uw] = v
        y.copy_(out.reshape(y.shape))

    # ---- GRU
    def gru_gate_fwd(self, pre_r, ld, h, rh, rows, Ch, flags=0, x_r=None):
        pre = _rows(pre_r, rows, ld, Ch)
        if x_r is not None:
            pre += _rows(x_r, rows, ld, Ch)     # completed in place (a view of pre_r)
        g = torch.sigmoid(pre)
        v = g * h.reshape(rows, Ch)
        if flags & 256:
            v = self._rna_tf32(v)
        rh.copy_(v.reshape(rh.shape))

    def gru_blend_fwd(self, pre_u, ld, h, c, hnew, hnew_tf32, rows, Ch, relu_c=False, x_u=None, x_c=None):
        pre = _rows(pre_u, rows, ld, Ch)
        if x_u is not None:
            pre += _rows(x_u, rows, ld, Ch)
        if x_c is not None:
            c += x_c.reshape(c.shape)
        u = torch.sigmoid(pre)
        cv = torch.relu(c.reshape(rows, Ch)) if relu_c else c.reshape(rows, Ch)
        v = u * h.reshape(rows, Ch) + (1 - u) * cv
        hnew.copy_(v.reshape(hnew.shape))
        if hnew_tf32 is not None:
            hnew_tf32.copy_(self._rna_tf32(v).reshape(hnew_tf32.shape))

    def gru_gate_bwd(self, d_rh, pre_r, ld, h, d_pre_r, ldd, dh, accumulate, rows, Ch, dz_scale=None, dz=None, dz_round=False):
        g = torch.sigmoid(_rows(pre_r, rows, ld, Ch))
        d = d_rh.reshape(rows, Ch)
        dp = d * h.reshape(rows, Ch) * g * (1 - g)
        _rows(d_pre_r, rows, ldd, Ch)[...] = dp
        if dz is not None:
            z = dp * dz_scale.reshape(1, Ch)
            _rows(dz, rows, ldd, Ch)[...] = self._rna_tf32(z) if dz_round else z
        v = (d * g).reshape(dh.shape)
        dh.add_(v) if accumulate else dh.copy_(v)

    def gru_blend_bwd(self, d_hnew, pre_u, ld, h, c, d_pre_u, ldd, dc, dh, accumulate, rows, Ch, relu_c=False, dz_u_scale=None, dz_u=None,
                      dz_c_scale=None, dz_c=None, dz_round=False):
        u = torch.sigmoid(_rows(pre_u, rows, ld, Ch))
        d = d_hnew.reshape(rows, Ch)
        cp = c.reshape(rows, Ch)
        cv = torch.relu(cp) if relu_c else cp
        dpu = d * (h.reshape(rows, Ch) - cv) * u * (1 - u)
        _rows(d_pre_u, rows, ldd, Ch)[...] = dpu
        if dz_u is not None:
            z = dpu * dz_u_scale.reshape(1, Ch)
            _rows(dz_u, rows, ldd, Ch)[...] = self._rna_tf32(z) if dz_round else z
        g = d * (1 - u)
        if relu_c:
            g = g * (cp > 0)
        dc.copy_(g.reshape(dc.shape))
        if dz_c is not None:
            z = g * dz_c_scale.reshape(1, Ch)
            dz_c.copy_((self._rna_tf32(z) if dz_round else z).reshape(dz_c.shape))
        v = (d * u).reshape(dh.shape)
        dh.add_(v) if accumulate else dh.copy_(v)

    # ---- BatchNorm
    def bn_stats(self, x, sums, rows, G, C):
        v = x.reshape(G, rows, C).double()
        sums.copy_(torch.stack([v.sum(1), (v * v).sum(1)], dim=-1))

    def bn_finalize(self, sums, gamma, beta, rmean, rvar, rows, G, C, eps, momentum, training, mean, invstd, a, b):
        for g in range(G):
            if training:
                m = sums[g, :, 0] / rows
                var = (sums[g, :, 1] / rows - m * m).clamp_min(0)
                unb = var * rows / (rows - 1) if rows > 1 else var
                rmean.copy_((1 - momentum) * rmean + momentum * m.float())
                rvar.copy_((1 - momentum) * rvar + momentum * unb.float())
                m, var = m.float(), var.float()
            else:
                m, var = rmean.clone(), rvar.clone()
            istd = 1.0 / torch.sqrt(var + eps)
            mean[g] = m
            invstd[g] = istd
            a[g] = gamma * istd
            b[g] = beta - m * a[g]

    @staticmethod
    def _bn_low(x, rows, G, C):
        return x.reshape(G, rows, C)

    def bn_apply(self, x, a, b, y, rows, G, C, relu, up2, H, W, x_rounded=None):
        rnd, relu = bool(int(relu) & 256), int(relu) & ~256
        if x_rounded is not None:
            assert not up2
            x_rounded.copy_(self._rna_tf32(x.reshape(x_rounded.shape)))
        v = x.reshape(G, rows, C) * a.reshape(G, 1, C) + b.reshape(G, 1, C)
        if relu:
            v = torch.relu(v)
        if rnd:
            v = self._rna_tf32(v)
        if up2:
            v = v.reshape(-1, H, W, C).repeat_interleave(2, 1).repeat_interleave(2, 2)
        y.copy_(v.reshape(y.shape))

    def _dpre(self, dy, x, a, b, rows, G, C, relu, up2, H, W):
        if up2:
            d = dy.reshape(-1, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(G, rows, C)
        else:
            d = dy.reshape(G, rows, C)
        if relu:
            yv = x.reshape(G, rows, C) * a.reshape(G, 1, C) + b.reshape(G, 1, C)
            d = d * (yv > 0)
        return d

    def bn_bwd_reduce(self, dy, x, a, b, mean, invstd, red, rows, G, C, relu, up2, H, W):
        d = self._dpre(dy, x, a, b, rows, G, C, relu, up2, H, W)
        xh = (x.reshape(G, rows, C) - mean.reshape(G, 1, C)) * invstd.reshape(G, 1, C)
        red.copy_(torch.stack([d.double().sum(1), (d * xh).double().sum(1)], dim=-1))

    def bn_bwd_apply(self, dy, x, a, b, mean, invstd, out_scale, red, dx, dgamma, dbeta, accumulate, rows, G, C, relu, up2, H, W, training,
                     dx_add=None):
        rnd, relu = bool(relu & 256), relu & ~256
        d = self._dpre(dy, x, a, b, rows, G, C, relu, up2, H, W)
        if dx is not None:
            if training:
                xh = (x.reshape(G, rows, C) - mean.reshape(G, 1, C)) * invstd.reshape(G, 1, C)
                m1 = (red[:, :, 0] / rows).float().reshape(G, 1, C)
                m2 = (red[:, :, 1] / rows).float().reshape(G, 1, C)
                v = a.reshape(G, 1, C) * (d - m1 - xh * m2)
            else:
                v = a.reshape(G, 1, C) * d
            if out_scale is not None:
                v = v * out_scale.reshape(G, 1, C)
            if dx_add is not None:
                v = v + dx_add.reshape(v.shape)
            if rnd:
                v = self._rna_tf32(v.contiguous())
            dx.copy_(v.reshape(dx.shape))
        if dgamma is not None:
            q = red[:, :, 1].sum(0).float()
            dgamma.add_(q) if accumulate else dgamma.copy_(q)
        if dbeta is not None:
            s = red[:, :, 0].sum(0).float()
            dbeta.add_(s) if accumulate else dbeta.copy_(s)

    # ---- spectral norm
    def sn_power_iter(self, w, u, v, R, K, G, eps, training, inv_sigma, u_hist, v_hist, ws):
        wm = w.reshape(R, K)
        for g in range(G):
            if training:
                t = torch.mv(wm, v)
                u.copy_(t / t.norm().clamp_min(eps))
                q = torch.mv(wm.t(), u)
                v.copy_(q / q.norm().clamp_min(eps))
            u_hist[g] = u
            v_hist[g] = v
            inv_sigma[g] = 1.0 / torch.dot(u, torch.mv(wm, v))

    def sn_power_iter_multi(self, items):
        for it in items:
            self.sn_power_iter(it["w"], it["u"], it["v"], it["R"], it["K"], it["G"], it["eps"], it["training"], it["inv_sigma"],
                               it["u_hist"], it["v_hist"], it["ws"])

    def sn_bwd(self, d_inv_sigma, inv_sigma, u_hist, v_hist, dw, R, K, G, accumulate):
        coef = -d_inv_sigma * inv_sigma * inv_sigma
        g = torch.einsum("g,gr,gk->rk", coef, u_hist, v_hist).reshape(dw.shape)
        dw.add_(g) if accumulate else dw.copy_(g)

    def rowdot_div(self, a, b, denom, out, rows, cols, ld, offset=0):
        av = a.reshape(rows, ld)[:, offset:offset + cols].double()
        bv = b.reshape(rows, ld)[:, offset:offset + cols].double()
        v = (av * bv).sum(1)
        if denom is not None:
            v = v / denom.reshape(-1)[:rows].double()
        out.reshape(-1)[:rows].copy_(v.float())

    def sn_bwd_multi(self, items):
        for it in items:
            self.sn_bwd(it["d_inv_sigma"], it["inv_sigma"], it["u_hist"], it["v_hist"], it["dw"], it["R"], it["K"], it["G"], it["accumulate"])

    # ---- conv
    def pack_weight(self, w, packed, Cout, CinTot, ci0, Cin, taps, mode):
        rnd, mode = bool(mode & 256), mode & ~256
        wv = w.reshape(Cout, CinTot, taps)[:, ci0:ci0 + Cin]
        if mode == 0:
            packed.copy_(wv.permute(2, 0, 1).reshape(-1))
        else:
            packed.copy_(wv.flip(2).permute(2, 1, 0).reshape(-1))
        if rnd:
            packed.copy_(self._rna_tf32(packed))

    def pack_weight_multi(self, items):
        for it in items:
            cout, cin, taps, pad, cot, co0 = it["Cout"], it["Cin"], it["taps"], it["CinPad"], it["CoutTot"], it["co0"]
            dense = torch.empty(taps * cout * cin)
            self.pack_weight(it["w"], dense, cout, it["CinTot"], it["ci0"], cin, taps, it["mode"])
            if (it["mode"] & ~256) == 0:
                it["packed"].view(taps, cot, pad)[:, co0:co0 + cout, :cin] = dense.view(taps, cout, cin)
            else:
                it["packed"].view(taps, pad, cot)[:, :cin, co0:co0 + cout] = dense.view(taps, cin, cout)

    def unpack_wgrad(self, packed, gw, Cout, CinTot, ci0, Cin, taps, accumulate):
        g = packed.reshape(taps, Cout, Cin).permute(1, 2, 0)
        tgt = gw.reshape(Cout, CinTot, taps)[:, ci0:ci0 + Cin]
        tgt.add_(g) if accumulate else tgt.copy_(g)

    @staticmethod
    def _conv_raw(x, wp, N, D, H, W, Cin, Cout, kd, kh, kw):
        w = wp.reshape(kd, kh, kw, Cout, Cin).permute(3, 4, 0, 1, 2)
        xi = x.reshape(N, D, H, W, Cin).permute(0, 4, 1, 2, 3)
        z = F.conv3d(xi, w, None, padding=(kd // 2, kh // 2, kw // 2))
        return z.permute(0, 2, 3, 4, 1)  # N,D,H,W,Cout

    def conv_fwd(self, x, wp, bias, scale, res, y, N, D, H, W, Cin, Cout, kd, kh, kw, G, act, algo=0, precision=0, x_lo=None, wp_lo=None):
        z = self._conv_raw(x, wp, N, D, H, W, Cin, Cout, kd, kh, kw)
        if scale is not None:
            z = (z.reshape(G, -1, Cout) * scale.reshape(G, 1, Cout)).reshape(N, D, H, W, Cout)
        if act & 512:  # DGMR_FLAG_ACCUMULATE
            y.add_(z.reshape(y.shape))
            return
        round_out, res_up2, act = bool(act & 1024), bool(act & 2048), act & 3
        if bias is not None:
            z = z + bias
        if res is not None:
            if res_up2:   # half-resolution residual, added nearest-upsampled
                r = res.reshape(N, D, H // 2, W // 2, Cout)
                r = r.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
                z = z + r
            else:
                z = z + res.reshape(z.shape)
        if act == 1:
            z = torch.relu(z)
        if round_out:
            z = self._rna_tf32(z.contiguous())
        y.copy_(z.reshape(y.shape))

    # ---- sub-pixel up-convolution (plain restatement of the phase formula of csrc/conv_subpix.cu)
    def upconv_supported(self, *a):
        return False

    @staticmethod
    def _tap_set(i, a):
        return ((0,), (1, 2))[a] if i == 0 else ((0, 1), (2,))[a]

    def pack_weight_subpix(self, w, packed, Cout, CinTot, ci0, Cin, mode):
        rnd, mode = bool(mode & 256), mode & ~256
        wv = w.reshape(Cout, CinTot, 3, 3)[:, ci0:ci0 + Cin]
        tiles = []
        for i in range(2):
            for j in range(2):
                for a in range(2):
                    for b in range(2):
                        t = sum(wv[:, :, kh, kw] for kh in self._tap_set(i, a) for kw in self._tap_set(j, b))
                        tiles.append(t if mode == 0 else t.t())
        p = torch.stack(tiles).reshape(-1)
        packed.copy_(self._rna_tf32(p.contiguous()) if rnd else p)

    def unpack_wgrad_subpix(self, dwsp, gw, Cout, CinTot, ci0, Cin, accumulate):
        t = dwsp.reshape(2, 2, 2, 2, Cout, Cin)
        g = torch.zeros(Cout, Cin, 3, 3)
        for i in range(2):
            for j in range(2):
                for a in range(2):
                    for b in range(2):
                        for kh in self._tap_set(i, a):
                            for kw in self._tap_set(j, b):
                                g[:, :, kh, kw] += t[i, j, a, b]
        tgt = gw.reshape(Cout, CinTot, 3, 3)[:, ci0:ci0 + Cin]
        tgt.add_(g) if accumulate else tgt.copy_(g)

    def _upconv_raw(self, x, wsp, N, H, W, Cin, Cout):
        xl = F.pad(x.reshape(N, H, W, Cin), (0, 0, 1, 1, 1, 1))     # zero halo
        t = wsp.reshape(2, 2, 2, 2, Cout, Cin)
        y = torch.zeros(N, 2 * H, 2 * W, Cout)
        for i in range(2):
            for j in range(2):
                acc = 0
                for a in range(2):
                    for b in range(2):
                        sl = xl[:, a + i:a + i + H, b + j:b + j + W]      # xl_unpadded[h + a + i - 1, w + b + j - 1]
                        acc = acc + torch.einsum("nhwc,oc->nhwo", sl, t[i, j, a, b])
                y[:, i::2, j::2] = acc
        return y

    def upconv_fwd(self, x, wsp, bias, scale, res, y, N, H, W, Cin, Cout, G, act):
        z = self._upconv_raw(x, wsp, N, H, W, Cin, Cout)
        if scale is not None:
            z = (z.reshape(G, -1, Cout) * scale.reshape(G, 1, Cout)).reshape(N, 2 * H, 2 * W, Cout)
        round_out, act = bool(act & 1024), act & 3
        if bias is not None:
            z = z + bias
        if res is not None:
            z = z + res.reshape(z.shape)
        if act == 1:
            z = torch.relu(z)
        if round_out:
            z = self._rna_tf32(z.contiguous())
        y.copy_(z.reshape(y.shape))

    def upconv_dgrad(self, dz, wspt, dx, N, H, W, Cin, Cout):
        wsp = wspt.reshape(16, Cin, Cout).transpose(1, 2).contiguous()
        with torch.enable_grad():
            x = torch.zeros(N, H, W, Cin, requires_grad=True)
            (g,) = torch.autograd.grad(self._upconv_raw(x, wsp, N, H, W, Cin, Cout), x, dz.reshape(N, 2 * H, 2 * W, Cout))
        dx.copy_(g.reshape(dx.shape))

    def upconv_wgrad(self, x, dz, dwsp, N, H, W, Cin, Cout):
        with torch.enable_grad():
            w = torch.zeros(16 * Cout * Cin, requires_grad=True)
            (g,) = torch.autograd.grad(self._upconv_raw(x.detach(), w, N, H, W, Cin, Cout), w, dz.reshape(N, 2 * H, 2 * W, Cout))
        dwsp.copy_(g)

    def conv_bwd_prep(self, dy, y, res, bias, scale, dz, dpre, dbias, dscale, rows, G, Cout, act, accumulate_dbias=False, up_hw=(0, 0), pool=None):
        rnd, act = bool(act & 256), act & ~256
        if pool:   # dy is the gradient of the average-pooled output: nearest-upsample it (zero on the floor-dropped rim), divide by the window
            pd, ph, pw, D, H, W = pool
            N = (rows * G) // (D * H * W)
            g = dy.reshape(N, D // pd, H // ph, W // pw, Cout) / float(pd * ph * pw)
            g = g.repeat_interleave(pd, 1).repeat_interleave(ph, 2).repeat_interleave(pw, 3)
            full = torch.zeros(N, D, H, W, Cout)
            full[:, :g.shape[1], :g.shape[2], :g.shape[3]] = g
            dy = full.reshape(G * rows, Cout)
        if res is not None and up_hw[0]:
            uh, uw = up_hw
            r = res.reshape(-1, uh // 2, uw // 2, Cout)
            res = r.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()
        d = dy.reshape(G, rows, Cout)
        if act == 1:
            d = d * (y.reshape(G, rows, Cout) > 0)
        if dz is not None:
            dz.copy_((d * scale.reshape(G, 1, Cout) if scale is not None else d).reshape(dz.shape))
            if rnd:
                dz.copy_(self._rna_tf32(dz))
        if dpre is not None:
            dpre.copy_(d.reshape(dpre.shape))
        if dbias is not None:
            s = d.sum(dim=(0, 1))
            dbias.add_(s) if accumulate_dbias else dbias.copy_(s)
        if dscale is not None:
            zs = y.reshape(G, rows, Cout)
            if bias is not None:
                zs = zs - bias
            if res is not None:
                zs = zs - res.reshape(G, rows, Cout)
            dscale.copy_((d * zs).sum(1) / scale.reshape(G, Cout))

    def conv_wgrad(self, x, dz, dwp, N, D, H, W, Cin, Cout, kd, kh, kw, algo=0, precision=0, x_lo=None, dz_lo=None):
        xi = x.reshape(N, D, H, W, Cin).permute(0, 4, 1, 2, 3)
        g = dz.reshape(N, D, H, W, Cout).permute(0, 4, 1, 2, 3)
        with torch.enable_grad():
            w = torch.zeros(Cout, Cin, kd, kh, kw, requires_grad=True)
            z = F.conv3d(xi.detach(), w, None, padding=(kd // 2, kh // 2, kw // 2))
            (gw,) = torch.autograd.grad(z, w, g.detach())
        dwp.copy_(gw.permute(2, 3, 4, 0, 1).reshape(-1))

    # ---- D head / attention / losses / optimiser
    def sumpool_relu_fwd(self, x, y, N, HW, C):
        y.copy_(torch.relu(x.reshape(N, HW, C)).sum(1))

    def sumpool_relu_bwd(self, dy, x, dx, N, HW, C):
        dx.copy_(((x.reshape(N, HW, C) > 0) * dy.reshape(N, 1, C)).reshape(dx.shape))

    @staticmethod
    def _att_mat(t, B, H, W, C):  # [B,1?,H,W,C] -> [B, C*H, W]
        return t.reshape(B, H, W, C).permute(0, 3, 1, 2).reshape(B, C * H, W)

    @staticmethod
    def _att_unmat(m, B, H, W, C):
        return m.reshape(B, C, H, W).permute(0, 2, 3, 1)

    def attention_fwd(self, q, k, v, out, beta, B, H, W, C):
        Q, K, V = (self._att_mat(t, B, H, W, C) for t in (q, k, v))
        be = torch.softmax(torch.bmm(Q, K.transpose(1, 2)), dim=-1)
        beta.copy_(be)
        out.copy_(self._att_unmat(torch.bmm(be, V), B, H, W, C).reshape(out.shape))

    def attention_bwd(self, dout, q, k, v, beta, dq, dk, dv, ws, B, H, W, C):
        Q, K, V, dO = (self._att_mat(t, B, H, W, C) for t in (q, k, v, dout))
        dV = torch.bmm(beta.transpose(1, 2), dO)
        dB = torch.bmm(dO, V.transpose(1, 2))
        dL = beta * (dB - (dB * beta).sum(-1, keepdim=True))
        ws.copy_(dL)
        dq.copy_(self._att_unmat(torch.bmm(dL, K), B, H, W, C).reshape(dq.shape))
        dk.copy_(self._att_unmat(torch.bmm(dL.transpose(1, 2), Q), B, H, W, C).reshape(dk.shape))
        dv.copy_(self._att_unmat(dV, B, H, W, C).reshape(dv.shape))

    def hinge_disc(self, scores, B, cols, loss, dscores):
        with torch.enable_grad():
            s = scores.detach().reshape(2 * B, cols).clone().requires_grad_(True)
            l = torch.relu(1 - s[:B]).mean(0).sum() + torch.relu(1 + s[B:]).mean(0).sum()
            (g,) = torch.autograd.grad(l, s)
        loss.copy_(l.detach())
        dscores.copy_(g.reshape(dscores.shape))

    def hinge_gen(self, scores, n, loss, dscores):
        loss.copy_(-scores.mean())
        dscores.fill_(-1.0 / n)

    def grid_cell_fwd(self, gen, target, cap, coef, loss, acc_ws):
        w = torch.clamp_min(target + 1, cap)
        loss.copy_((((gen - target) * w).abs().double().sum() * coef).float())

    def grid_cell_bwd(self, gen, target, cap, coef, gout, dgen):
        w = torch.clamp_min(target + 1, cap)
        dgen.copy_(torch.sign((gen - target) * w) * w * coef * gout)

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
        gr = g * grad_scale
        m.mul_(beta1).add_(gr, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
