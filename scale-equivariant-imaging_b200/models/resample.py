"""The "ideal" (Fourier-domain) resamplers of the reference's CNN as explicit linear operators.

Reference: src/models/convolutional.py, IdealUpsample.forward (:54-92) and IdealDownsample.forward (:113-133):

    rfft2 -> fftshift over BOTH axes (also the half-spectrum one) -> zero-pad (up) / mask (down)
          -> [ifftshift result discarded] -> irfft2 (-> [::rate, ::rate] for down)

Every step is linear and acts on one axis at a time, so the chain is a fixed linear map of the image.  Along the
height the map is complex (fft, shift = modulation, mask, ifft): G = Gr + i Gi.  Along the width it goes real ->
half spectrum R = Rr + i Ri -> real through the real-linear c2r transform (Ir, Ii).  Expanding the complex product,

    out = Gr X P^T + Gi X Q^T,     P = Ir Rr + Ii Ri,   Q = Ii Rr - Ir Ri.

Gi does not vanish because the kept frequency band [hc, H - hc) of the shifted spectrum is not symmetric.  The four
matrices are obtained by pushing identity matrices through torch.fft in float64, i.e. they ARE the reference's
operator (tests/test_cnn_kernels.py::test_resample_operators_match_reference: 1e-12 against fixtures from the reference).

On the GPU the two contractions run as batched tensor-core products on channels-last bf16 activations
(csrc/bgemm.cu, sei_bgemm_bf16): width first over batch = (image, row) with A = [P; Q], then height over
batch = image with A = [Gr | Gi] (columns interleaved to match the (row, term) order of the intermediate); the backward
pass applies the transposed operators the same way.  Small operators are applied to several batch entries at once through
a block-diagonal operator (_Product).
"""
from math import ceil

import torch

from sei_b200 import ops

_OPERATORS = {}
_PACKED = {}


def operator(kind, H, W, rate):
    """float64 CPU matrices (Gr, Gi, P, Q) with out = Gr @ X @ P.T + Gi @ X @ Q.T for one (H, W) image"""
    key = (kind, H, W, rate)
    if key in _OPERATORS:
        return _OPERATORS[key]
    c128 = torch.complex128
    Fh = torch.fft.fftshift(torch.fft.fft(torch.eye(H, dtype=c128), dim=0), dim=0)                   # (H, H)
    Rw = torch.fft.fftshift(torch.fft.rfft(torch.eye(W, dtype=torch.float64), dim=0), dim=0)         # (W/2+1, W)
    ws = W // 2 + 1
    if kind == "down":
        hcsh, hcsw = ceil(H / (2 * rate)), ceil(ws / (2 * rate))
        mh = torch.zeros(H, dtype=torch.float64)
        mh[hcsh:H - hcsh] = 1                      # otf[:, :, hcsh:-hcsh, hcsw:-hcsw] = 1  (:124-125)
        mw = torch.zeros(ws, dtype=torch.float64)
        mw[hcsw:ws - hcsw] = 1
        G = torch.fft.ifft(mh[:, None] * Fh, dim=0)
        Rw = mw[:, None] * Rw
        n_out = W
    elif kind == "up":
        mv, mhz = (H * (rate - 1)) // 2, (ws * (rate - 1)) // 2
        mt = mv + 1 if H % 2 == 1 else mv         # margins (:66-82)
        ml = mhz + 1 if ws % 2 == 1 else mhz
        Fp = torch.zeros((H * rate, H), dtype=c128)
        Fp[mt:mt + H] = Fh
        G = torch.fft.ifft(Fp, dim=0)
        n_out = W * rate
        nh = n_out // 2 + 1                        # irfft reads the first n/2+1 columns of the padded half spectrum
        Rp = torch.zeros((max(ws * rate, nh), W), dtype=c128)
        Rp[ml:ml + ws] = Rw
        Rw = Rp[:nh]
    else:
        raise ValueError(kind)
    E = torch.eye(Rw.shape[0], dtype=c128)
    Ir = torch.fft.irfft(E, n=n_out, dim=0)
    Ii = torch.fft.irfft(1j * E, n=n_out, dim=0)
    P = Ir @ Rw.real + Ii @ Rw.imag
    Q = Ii @ Rw.real - Ir @ Rw.imag
    Gr, Gi = G.real, G.imag
    if kind == "down":
        Gr, Gi, P, Q = Gr[::rate], Gi[::rate], P[::rate], Q[::rate]     # out[:, :, ::rate, ::rate]  (:133)
    _OPERATORS[key] = tuple(m.contiguous() for m in (Gr, Gi, P, Q))
    return _OPERATORS[key]


def apply_dense(kind, x, rate):
    """the operator applied with plain matmuls in x's dtype (reference formulation for tests; any device)"""
    Gr, Gi, P, Q = (m.to(device=x.device, dtype=x.dtype) for m in operator(kind, x.shape[-2], x.shape[-1], rate))
    return Gr @ x @ P.t() + Gi @ x @ Q.t()


class _Packed:
    """one operator matrix A[M, K] in the layout sei_bgemm_bf16 wants: bf16, K padded to 64, rows to the CTA tile"""

    def __init__(self, A, device, dtype=torch.bfloat16):
        self.M, self.K = A.shape
        kpad = -(-self.K // 64) * 64
        self.tile = ops.bgemm_tile_rows(self.M, kpad)
        if self.tile <= 0:
            raise ops.SeiError(f"resample operator {self.M}x{self.K} does not fit the shared-memory resident tile")
        rows = -(-self.M // self.tile) * self.tile
        buf = torch.zeros((rows, kpad), dtype=torch.float64)
        buf[:self.M, :self.K] = A
        self.data = buf.to(device=device, dtype=dtype)


class _Product:
    """One of the four batched products of a resampler: D_b = A @ X_b with contiguous items (X_b: [K, N], D_b: [M, N],
    consecutive batch entries back to back).  A small operator (M < 128 rows) fills only part of the 128 TMEM lanes of a
    UMMA and leaves the epilogue of the tcgen05 kernel to one or two of its four warps (measured: 0.33-0.44 of the copy
    peak at the deepest level, profiles/r01_resample_bench.md), so P consecutive batch entries are processed as ONE item
    with the block-diagonal operator I_P (x) A: [P M, P K] @ [P K, N] -- the stacked entries ARE a contiguous [P K, N]
    matrix.  The extra zero blocks cost tensor-core time only, of which there is plenty (the product is HBM-bound)."""

    def __init__(self, A, device, dtype=torch.bfloat16):
        self.A, self.device, self.dtype, self._packs = A, device, dtype, {}

    def _pack_factor(self, batches):
        M, K = self.A.shape
        pow2 = lambda v: v > 0 and (v & (v - 1)) == 0          # noqa: E731
        if not (pow2(M) and pow2(K)) or M >= 128:
            return 1
        P = 1
        while P < 8 and 2 * P * M <= 128 and 2 * P * K <= 512 and batches % (2 * P) == 0:
            P *= 2
        return P

    def __call__(self, src, dst, N, batches):
        P = self._pack_factor(batches)
        a = self._packs.get(P)
        if a is None:
            a = self._packs[P] = _Packed(torch.block_diag(*([self.A] * P)) if P > 1 else self.A, self.device, self.dtype)
        ops.bgemm_bf16(a.data, src, dst, a.M, a.K, N, a.tile, batches // P, 1, (a.K * N, 0), a.K, (0, N),
                       (a.M * N, 0), a.M, (0, N))
        return dst


def _packed(kind, H, W, rate, device, dtype=torch.bfloat16):
    """the four products of one resampler; dtype other than bf16 only for the CPU emulation in tests/test_cnn_kernels.py"""
    key = (kind, H, W, rate, str(device), str(dtype))
    if key not in _PACKED:
        Gr, Gi, P, Q = operator(kind, H, W, rate)
        Ho, Wo = Gr.shape[0], P.shape[0]
        A1 = torch.cat([P, Q], 0)                                             # (2 Wo, W): rows (term, wo)
        A2 = torch.stack([Gr, Gi], 2).reshape(Ho, 2 * H)                      # (Ho, 2 H): columns (h, term) interleaved
        _PACKED[key] = {
            "Ho": Ho, "Wo": Wo,
            "A1": _Product(A1, device, dtype), "A2": _Product(A2, device, dtype),
            "A2T": _Product(A2.t().contiguous(), device, dtype), "A1T": _Product(A1.t().contiguous(), device, dtype),
        }
    return _PACKED[key]


def _forward_cl(x, pk):
    """x: (B, H, W, C) contiguous bf16 -> (B, Ho, Wo, C).  The two-term intermediate is laid out (B, H, 2, Wo, C): every
    product then reads and writes plain contiguous matrices -- per image row [W, C] -> [2 Wo, C] along the width, per image
    [2 H, Wo C] -> [Ho, Wo C] along the height (rows (h, term) match the interleaved columns of A2)."""
    B, H, W, C = x.shape
    Ho, Wo = pk["Ho"], pk["Wo"]
    y = torch.empty((B, H, 2, Wo, C), dtype=x.dtype, device=x.device)
    pk["A1"](x, y, C, B * H)
    out = torch.empty((B, Ho, Wo, C), dtype=x.dtype, device=x.device)
    return pk["A2"](y, out, Wo * C, B)


def _backward_cl(g, pk, H, W):
    """transposed operator: g (B, Ho, Wo, C) -> (B, H, W, C)"""
    B, Ho, Wo, C = g.shape
    gy = torch.empty((B, H, 2, Wo, C), dtype=g.dtype, device=g.device)
    pk["A2T"](g, gy, Wo * C, B)
    gx = torch.empty((B, H, W, C), dtype=g.dtype, device=g.device)
    return pk["A1T"](gy, gx, C, B * H)


class _IdealResample(torch.autograd.Function):
    """x: logical (B, C, H, W), channels-last memory, bf16, C % 8 == 0"""

    @staticmethod
    def forward(ctx, x, kind, rate):
        B, C, H, W = x.shape
        pk = _packed(kind, H, W, rate, x.device)
        ctx.pk, ctx.hw = pk, (H, W)
        xl = x.permute(0, 2, 3, 1).contiguous()            # a view for channels-last tensors
        return _forward_cl(xl, pk).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        gl = g.permute(0, 2, 3, 1).contiguous()
        return _backward_cl(gl, ctx.pk, *ctx.hw).permute(0, 3, 1, 2), None, None


_CONST = {}


def constant_response(kind, H, W, rate, device):
    """the operator applied to the all-ones image, (Ho, Wo) fp32: what a per-channel constant (a bias) turns into.
    It is NOT constant: the reference's fftshift without the matching ifftshift moves DC to a high frequency."""
    key = (kind, H, W, rate, str(device))
    if key not in _CONST:
        Gr, Gi, P, Q = operator(kind, H, W, rate)
        pat = torch.outer(Gr.sum(1), P.sum(1)) + torch.outer(Gi.sum(1), Q.sum(1))
        _CONST[key] = pat.to(device=device, dtype=torch.float32)
    return _CONST[key]


def supported(x):
    return x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0


def ideal_resample(x, kind, rate):
    return _IdealResample.apply(x, kind, rate)
