"""ctypes binding of libmfpa.so (include/mfpa.h).

There is no CPU fallback: importing this module raises ``ImportError`` when the
shared library has not been built (``python -m musicfpaugment_b200.build``),
and ``Context()`` raises when no sm_100 GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFPA_LIB", os.path.join(_HERE, "libmfpa.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing - build it with `python -m musicfpaugment_b200.build`; "
        "musicfpaugment_b200 has no CPU fallback"
    )

_lib = C.CDLL(LIB_PATH)


class AfpParams(C.Structure):
    """mfpa_afp_params (include/mfpa.h)."""

    _fields_ = [
        ("a_dec", C.c_double), ("f_sd", C.c_double), ("maxpks", C.c_int32), ("mindt", C.c_int32),
        ("targetdt", C.c_int32), ("targetdf", C.c_int32), ("fanout", C.c_int32), ("reserved", C.c_int32),
    ]


class AugParams(C.Structure):
    """mfpa_aug_params (include/mfpa.h): one record per query."""

    _fields_ = [
        ("apply", C.c_uint32), ("fc1_hz", C.c_float), ("fc2_hz", C.c_float), ("fc3_hz", C.c_float),
        ("snr_db", C.c_float), ("gain_factor", C.c_float), ("clip_p", C.c_float), ("ir_len", C.c_int32),
    ]


AUG_HPF1, AUG_IR, AUG_NOISE, AUG_GAIN, AUG_CLIP, AUG_LPF, AUG_HPF3, AUG_NORM = 1, 2, 4, 8, 16, 32, 64, 128
AUG_ALL = 255
AUG_DTYPE = [("apply", "<u4"), ("fc1_hz", "<f4"), ("fc2_hz", "<f4"), ("fc3_hz", "<f4"), ("snr_db", "<f4"),
             ("gain_factor", "<f4"), ("clip_p", "<f4"), ("ir_len", "<i4")]
# mfpa_noise_piece
NOISE_PIECE_DTYPE = [("src_a", "<i8"), ("src_b", "<i8"), ("query", "<i4"), ("dst", "<i4"), ("len", "<i4"), ("reserved", "<i4")]


class ChainInputs(C.Structure):
    """mfpa_chain_inputs (include/mfpa.h)."""

    _fields_ = [("x_host", C.c_void_p), ("x_pcm16_host", C.c_void_p), ("ir_dev", C.c_void_p), ("ir_bank_len", C.c_int64),
                ("ir_offsets_host", C.c_void_p), ("ir_stride", C.c_int32), ("n_pieces", C.c_int32), ("noise_dev", C.c_void_p),
                ("noise_bank_dev", C.c_void_p), ("noise_bank_len", C.c_int64), ("pieces_host", C.c_void_p)]


class MatchParams(C.Structure):
    """mfpa_match_params (include/mfpa.h) = Matcher.__init__ defaults."""

    _fields_ = [("window", C.c_int32), ("threshcount", C.c_int32), ("search_depth", C.c_int32),
                ("max_alignments_per_id", C.c_int32)]


MAX_PEERS = 16


class PeerSet(C.Structure):
    """mfpa_peer_set (include/mfpa.h): every rank's receive buffers of the fused sparse exchange, as device pointers
    valid in this process (the rank's own allocation, the others mapped through CUDA IPC)."""

    _fields_ = [("words", C.c_void_p * MAX_PEERS), ("nwords", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS),
                ("world", C.c_int32), ("rank", C.c_int32)]


class MfpaError(RuntimeError):
    pass


_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
_P = C.POINTER(AfpParams)

# name -> (restype, argtypes); every symbol declared in include/mfpa.h
SIGNATURES = {
    "mfpa_abi_version": (_i, []),
    "mfpa_last_error": (C.c_char_p, []),
    "mfpa_create": (_i, [C.POINTER(_vp), _i]),
    "mfpa_destroy": (None, [_vp]),
    "mfpa_afp_defaults": (None, [_P]),
    "mfpa_set_spread_table": (_i, [_vp, _vp]),
    "mfpa_set_option": (_i, [_vp, _i, _i]),
    "mfpa_stage_times": (_i, [_vp, _vp]),
    "mfpa_num_frames": (_i, [_i]),
    "mfpa_shift_offset": (_i, [_i, _i]),
    "mfpa_stft_mag": (_i, [_vp, _vp, _i, _i, _i64, _i, _vp, _vp, _vp]),
    "mfpa_stft_num_frames": (_i, [_i, _i, _i, _i]),
    "mfpa_stft_complex": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "mfpa_spec_from_mag": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "mfpa_audfprint_peaks": (_i, [_vp, _vp, _vp, _i, _i, _i, _P, _vp, _vp, _vp]),
    "mfpa_audfprint_peaks_from_spec": (_i, [_vp, _vp, _i, _i, _i, _P, _vp, _vp, _vp]),
    "mfpa_peaks_list": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "mfpa_peaks_mask": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "mfpa_landmark_hashes": (_i, [_vp, _vp, _i, _i, _P, _i, _vp, _i, _vp, _vp]),
    "mfpa_merge_shifts": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "mfpa_fingerprint": (_i, [_vp, _vp, _i, _i, _i64, _i, _P, _vp, _i, _vp, _vp]),
    "mfpa_fingerprint_host": (_i, [_vp, _vp, _i, _i, _i, _P, _vp, _i64, _vp]),
    "mfpa_fingerprint_host_pcm16": (_i, [_vp, _vp, _i, _i, _i, _P, _vp, _i64, _vp]),
    "mfpa_augment": (_i, [_vp, _vp, _i, _i, _i64, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "mfpa_noise_assemble": (_i, [_vp, _vp, _i64, _vp, _i, _i, _i, _vp, _vp]),
    "mfpa_augment_fingerprint": (_i, [_vp, _vp, _i, _i, _i64, _i, _vp, _vp, _i, _vp, _i, _P, _vp, _i, _vp, _vp]),
    "mfpa_augment_fingerprint_host": (_i, [_vp, C.POINTER(ChainInputs), _i, _i, _i, _vp, _i, _P, _vp, _i64, _vp]),
    "mfpa_lowpass_filters": (_i, [_vp, _vp, _i, _i, _i64, _vp, _vp, _vp, _vp]),
    "mfpa_index_load": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i]),
    "mfpa_match_defaults": (None, [C.POINTER(MatchParams)]),
    "mfpa_get_hits": (_i, [_vp, _vp, _i, _vp, _i64, _vp, _vp]),
    "mfpa_match_counts": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "mfpa_match_select": (_i, [_vp, _vp, _i, C.POINTER(MatchParams), _vp, _vp, _vp]),
    "mfpa_match_collect": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, C.POINTER(MatchParams), _vp, _i, _vp, _vp]),
    "mfpa_match_align": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, C.POINTER(MatchParams), _vp, _vp, _i, _vp]),
    "mfpa_match_emit": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "mfpa_match_owner": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(MatchParams), _vp, _vp, _i, _vp]),
    "mfpa_match": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(MatchParams), _vp, _vp, _i, _vp]),
    "mfpa_peer_alloc": (_i, [_vp, C.c_uint64, C.POINTER(_vp), _vp]),
    "mfpa_peer_open": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "mfpa_peer_close": (_i, [_vp, _vp]),
    "mfpa_peer_free": (_i, [_vp, _vp]),
    "mfpa_match_emit_peer": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(PeerSet), _i, _vp]),
    "mfpa_peer_barrier": (_i, [_vp, C.POINTER(PeerSet), C.c_uint32, _vp]),
    "mfpa_dejavu_peaks": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.c_double, _vp, _vp, _i, _vp, _vp]),
    "mfpa_dejavu_num_frames": (_i, [_i]),
    "mfpa_dejavu_psd": (_i, [_vp, _vp, _i, _i, _i64, _vp, _vp]),
    "mfpa_dejavu_log": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mfpa_dejavu_index_create": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, C.POINTER(_vp)]),
    "mfpa_dejavu_index_destroy": (None, [_vp]),
    "mfpa_dejavu_return_matches": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i64, _vp, _vp, _vp]),
    "mfpa_dejavu_align": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "mfpa_mask_metrics": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "mfpa_psnr_stats": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "mfpa_compact_rows": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i64, _vp]),
    "mfpa_unet_num_params": (_i64, []),
    "mfpa_unet_create": (_i, [_vp, C.POINTER(_vp)]),
    "mfpa_unet_destroy": (None, [_vp]),
    "mfpa_unet_load": (_i, [_vp, _vp, _i64]),
    "mfpa_unet_set_max_chunk": (_i, [_vp, _i]),
    "mfpa_unet_forward": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i, _i, _i, _vp, _i64, _i64, _i64, _vp]),
    "mfpa_conv_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
}


def _bind(sigs):
    for name, (res, args) in sigs.items():
        fn = getattr(_lib, name)
        fn.restype = res
        fn.argtypes = args


_bind(SIGNATURES)
ALL_SIGNATURES = dict(SIGNATURES)

OPT_PEAKS_F64 = 1
OPT_MATCH_PACKED = 2
OPT_MATCH_UNFUSED = 3
OPT_PART_BUDGET_MB = 4
OPT_CONV_OCC = 5
OPT_STAGE_TIMES = 6
OPT_CLIP_POOLED = 7
STAGE_NAMES = ("hpf1_filter", "hpf1_conv", "ir_filter", "ir_conv", "mix", "clip_lpf", "hpf3_filter", "hpf3_conv", "stft",
               "peaks", "landmarks")
N_FFT, HOP, BINS, ROWS, MAG_PITCH, MAX_PKS, MAX_SHIFTS, HASHES_PER_FRAME = 512, 256, 257, 256, 264, 5, 8, 15


def last_error() -> str:
    return _lib.mfpa_last_error().decode()


def check(rc: int) -> None:
    if rc != 0:
        raise MfpaError(f"libmfpa error {rc}: {last_error()}")


def raw():
    """The ctypes CDLL handle (for code that needs an entry point directly)."""
    return _lib


def num_frames(n_samples: int) -> int:
    return _lib.mfpa_num_frames(n_samples)


def shift_offset(shift: int, shifts: int) -> int:
    return _lib.mfpa_shift_offset(shift, shifts)


def afp_defaults() -> AfpParams:
    p = AfpParams()
    _lib.mfpa_afp_defaults(C.byref(p))
    return p


def match_defaults() -> MatchParams:
    p = MatchParams()
    _lib.mfpa_match_defaults(C.byref(p))
    return p


def _ptr(t):
    if isinstance(t, int):   # a raw device address (peer-mapped memory has no tensor around it)
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _row_stride(x) -> int:
    """Row stride in elements; a size-1 batch dimension may carry a meaningless stride."""
    return int(x.stride(0)) if x.shape[0] > 1 else int(x.shape[1])


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """Owns one ``mfpa_ctx`` on one GPU.  Methods take/return torch CUDA tensors;
    torch is used for device memory and streams only."""

    def __init__(self, device: int | None = None, spread_table=None):
        import torch

        if not torch.cuda.is_available():
            raise MfpaError("no CUDA device visible: musicfpaugment_b200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self._h = C.c_void_p()
        check(_lib.mfpa_create(C.byref(self._h), self.device))
        if spread_table is not None:
            self.set_spread_table(spread_table)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.mfpa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ---- helpers -------------------------------------------------------
    def _dev(self):
        import torch

        return torch.device("cuda", self.device)

    def set_spread_table(self, table):
        import numpy as np

        t = np.ascontiguousarray(table, dtype=np.float64)
        assert t.shape == (2 * ROWS + 1,)
        check(_lib.mfpa_set_spread_table(self._h, t.ctypes.data_as(C.c_void_p)))

    def set_option(self, option: int, value: int):
        check(_lib.mfpa_set_option(self._h, option, value))
        if option == OPT_MATCH_PACKED:
            self.packed_counts = bool(value)

    def stage_times(self):
        """Mean milliseconds per stage over the last (up to 16) chain / fingerprint calls made with OPT_STAGE_TIMES set:
        ({stage name: ms}, number of calls averaged)."""
        ms = (C.c_float * len(STAGE_NAMES))()
        n = _lib.mfpa_stage_times(self._h, ms)
        if n < 0:
            check(n)
        return {k: float(v) for k, v in zip(STAGE_NAMES, ms)}, n

    # ---- S2 --------------------------------------------------------------
    def stft_mag(self, x, shifts: int = 1):
        """x: [B,T] float32 cuda -> (mag [B*shifts, N, 264] f32, qmax [B*shifts] f32)."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
        B, T = x.shape
        n = num_frames(T)
        mag = torch.empty(B * shifts, n, MAG_PITCH, dtype=torch.float32, device=x.device)
        qmax = torch.empty(B * shifts, dtype=torch.float32, device=x.device)
        check(_lib.mfpa_stft_mag(self._h, _ptr(x), B, T, _row_stride(x), shifts, _ptr(mag), _ptr(qmax), _stream()))
        return mag, qmax

    def stft_complex(self, x, n_fft: int, hop: int, window):
        """stft.stft (afp/audfprint/stft.py:15-62): x float64 [T] cuda, window float64 [L] cuda ->
        complex128 [n_fft/2+1, frames] cuda."""
        import torch

        assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 1 and x.is_contiguous()
        assert window.is_cuda and window.dtype == torch.float64 and window.dim() == 1 and window.is_contiguous()
        n = _lib.mfpa_stft_num_frames(x.numel(), n_fft, hop, window.numel())
        out = torch.empty(n_fft // 2 + 1, max(n, 0), dtype=torch.complex128, device=x.device)
        check(_lib.mfpa_stft_complex(self._h, _ptr(x), x.numel(), n_fft, hop, _ptr(window), window.numel(), _ptr(out), _stream()))
        return out

    def spec_from_mag(self, mag, qmax, T: int, shifts: int = 1):
        import torch

        items, n = mag.shape[0], mag.shape[1]
        spec = torch.empty(items, BINS, n, dtype=torch.float64, device=mag.device)
        check(_lib.mfpa_spec_from_mag(self._h, _ptr(mag), _ptr(qmax), items // shifts, T, shifts, _ptr(spec), _stream()))
        return spec

    # ---- S3 --------------------------------------------------------------
    def audfprint_peaks(self, mag, qmax, T: int, shifts: int, params: AfpParams):
        """-> (rec uint64-as-int64 [items, N], npeaks int32 [items])."""
        import torch

        items, n = mag.shape[0], mag.shape[1]
        rec = torch.empty(items, n, dtype=torch.int64, device=mag.device)
        npk = torch.empty(items, dtype=torch.int32, device=mag.device)
        check(_lib.mfpa_audfprint_peaks(self._h, _ptr(mag), _ptr(qmax), items // shifts, T, shifts,
                                        C.byref(params), _ptr(rec), _ptr(npk), _stream()))
        return rec, npk

    def audfprint_peaks_from_spec(self, spec, stage: int, params: AfpParams):
        """spec: [items, 257|256, N] float64 cuda in the reference layout."""
        import torch

        assert spec.is_cuda and spec.dtype == torch.float64 and spec.dim() == 3
        spec = spec.contiguous()
        items, rows, n = spec.shape
        assert rows == (BINS if stage == 0 else ROWS)
        rec = torch.empty(items, n, dtype=torch.int64, device=spec.device)
        npk = torch.empty(items, dtype=torch.int32, device=spec.device)
        check(_lib.mfpa_audfprint_peaks_from_spec(self._h, _ptr(spec), items, n, stage, C.byref(params),
                                                  _ptr(rec), _ptr(npk), _stream()))
        return rec, npk

    def peaks_list(self, rec):
        import torch

        items, n = rec.shape
        cap = MAX_PKS * n
        peaks = torch.empty(items, cap, 2, dtype=torch.int32, device=rec.device)
        npk = torch.empty(items, dtype=torch.int32, device=rec.device)
        check(_lib.mfpa_peaks_list(self._h, _ptr(rec), items, n, _ptr(peaks), cap, _ptr(npk), _stream()))
        return peaks, npk

    def peaks_mask(self, rec):
        import torch

        items, n = rec.shape
        mask = torch.empty(items, ROWS, n, dtype=torch.float32, device=rec.device)
        check(_lib.mfpa_peaks_mask(self._h, _ptr(rec), items, n, _ptr(mask), _stream()))
        return mask

    # ---- S4 --------------------------------------------------------------
    def landmark_hashes(self, rec, params: AfpParams, sorted_rows: bool = False):
        import torch

        items, n = rec.shape
        cap = HASHES_PER_FRAME * n
        hashes = torch.empty(items, cap, 2, dtype=torch.int32, device=rec.device)
        nh = torch.empty(items, dtype=torch.int32, device=rec.device)
        check(_lib.mfpa_landmark_hashes(self._h, _ptr(rec), items, n, C.byref(params), int(sorted_rows),
                                        _ptr(hashes), cap, _ptr(nh), _stream()))
        return hashes, nh

    def merge_shifts(self, hashes, nh, shifts: int, n_frames: int):
        import torch

        items, cap_in, _ = hashes.shape
        B = items // shifts
        cap_out = cap_in * shifts
        out = torch.empty(B, cap_out, 2, dtype=torch.int32, device=hashes.device)
        nout = torch.empty(B, dtype=torch.int32, device=hashes.device)
        check(_lib.mfpa_merge_shifts(self._h, _ptr(hashes), _ptr(nh), B, shifts, cap_in, n_frames, _ptr(out),
                                     cap_out, _ptr(nout), _stream()))
        return out, nout

    # ---- S1 --------------------------------------------------------------
    @staticmethod
    def _aug_args(x, params, ir, noise):
        import numpy as np
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
        params = np.ascontiguousarray(params, dtype=AUG_DTYPE)
        assert params.shape == (x.shape[0],)
        if ir is not None:
            assert ir.is_cuda and ir.dtype == torch.float32 and ir.dim() == 2 and ir.is_contiguous()
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and noise.shape == x.shape
        return params

    def augment(self, x, params, ir=None, noise=None, sample_rate: int = 8000):
        """AugmentFP chain on dumped parameters.  x [B,T] f32 cuda; params: numpy structured
        array (AUG_DTYPE) [B]; ir [B,Lmax] f32 cuda; noise [B,T] f32 cuda -> [B,T] f32 cuda."""
        import torch

        params = self._aug_args(x, params, ir, noise)
        B, T = x.shape
        out = torch.empty(B, T, dtype=torch.float32, device=x.device)
        check(_lib.mfpa_augment(self._h, _ptr(x), B, T, _row_stride(x), sample_rate, params.ctypes.data_as(C.c_void_p),
                                _ptr(ir), ir.shape[1] if ir is not None else 0, _ptr(noise), _ptr(out), _stream()))
        return out

    def lowpass_filters(self, x, cutoff, width=None):
        """One julius.LowPassFilters filter per row: x [B,T] f32 cuda, cutoff / width float64 [B] fractions of the sample
        rate (width = the cut-off whose window the bank uses; default the filter's own) -> [B,T] f32 cuda."""
        import numpy as np
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
        B, T = x.shape
        cutoff = np.ascontiguousarray(cutoff, dtype=np.float64).reshape(B)
        width = cutoff if width is None else np.ascontiguousarray(width, dtype=np.float64).reshape(B)
        out = torch.empty(B, T, dtype=torch.float32, device=x.device)
        check(_lib.mfpa_lowpass_filters(self._h, _ptr(x), B, T, _row_stride(x), cutoff.ctypes.data_as(C.c_void_p),
                                        width.ctypes.data_as(C.c_void_p), _ptr(out), _stream()))
        return out

    def noise_assemble(self, bank, pieces, B: int, T: int):
        """AddBackgroundNoise.random_background for a batch (background_noise.py:64-141) from a device bank:
        bank f32 [bank_len] cuda; pieces: numpy structured array (NOISE_PIECE_DTYPE) -> [B,T] f32 cuda."""
        import numpy as np
        import torch

        assert bank.is_cuda and bank.dtype == torch.float32 and bank.dim() == 1 and bank.is_contiguous()
        pieces = np.ascontiguousarray(pieces, dtype=NOISE_PIECE_DTYPE)
        out = torch.empty(B, T, dtype=torch.float32, device=bank.device)
        check(_lib.mfpa_noise_assemble(self._h, _ptr(bank), bank.numel(), pieces.ctypes.data_as(C.c_void_p), len(pieces),
                                       B, T, _ptr(out), _stream()))
        return out

    def augment_fingerprint(self, x, params, ir, noise, shifts: int, afp: AfpParams, sample_rate: int = 8000):
        """Fused S1-S4: degraded-query hashes without leaving the device."""
        import torch

        params = self._aug_args(x, params, ir, noise)
        B, T = x.shape
        cap = HASHES_PER_FRAME * num_frames(T) * shifts
        out = torch.empty(B, cap, 2, dtype=torch.int32, device=x.device)
        nh = torch.empty(B, dtype=torch.int32, device=x.device)
        check(_lib.mfpa_augment_fingerprint(self._h, _ptr(x), B, T, _row_stride(x), sample_rate,
                                            params.ctypes.data_as(C.c_void_p), _ptr(ir),
                                            ir.shape[1] if ir is not None else 0, _ptr(noise), shifts, C.byref(afp),
                                            _ptr(out), cap, _ptr(nh), _stream()))
        return out, nh

    def augment_fingerprint_host(self, x, params, shifts: int, afp: AfpParams, ir=None, ir_offsets=None, noise=None,
                                 noise_bank=None, pieces=None, sample_rate: int = 8000, rows=None, offsets=None):
        """AugmentFP chain + wavfile2hashes for a batch of HOST queries (x: [B,T] float32 or int16 PCM, CPU torch tensor
        or numpy; pinned memory overlaps the copies with the kernels).  Degradation sources are device tensors: `ir`
        [B,L] rows (or a 1-D bank with `ir_offsets` int64 [B]), `noise` [B,T] rows (or `noise_bank` 1-D + `pieces`,
        a NOISE_PIECE_DTYPE array grouped by query).  -> (rows int32 [total,2], offsets int64 [B+1]) numpy views."""
        import numpy as np
        import torch

        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int16 if x.dtype == np.int16 else np.float32))
        assert not x.is_cuda and x.dtype in (torch.float32, torch.int16) and x.dim() == 2 and x.is_contiguous()
        B, T = x.shape
        params = np.ascontiguousarray(params, dtype=AUG_DTYPE)
        assert params.shape == (B,)
        inp = ChainInputs()
        if x.dtype == torch.int16:
            inp.x_pcm16_host = x.data_ptr()
        else:
            inp.x_host = x.data_ptr()
        keep = [x, params]
        if ir is not None:
            assert ir.is_cuda and ir.dtype == torch.float32 and ir.is_contiguous()
            inp.ir_dev = ir.data_ptr()
            if ir_offsets is not None:
                offs = np.ascontiguousarray(ir_offsets, dtype=np.int64)
                assert offs.shape == (B,) and ir.dim() == 1
                inp.ir_offsets_host, inp.ir_bank_len, inp.ir_stride = offs.ctypes.data, ir.numel(), 0
                keep.append(offs)
            else:
                assert ir.dim() == 2 and ir.shape[0] == B
                inp.ir_stride = ir.shape[1]
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and tuple(noise.shape) == (B, T)
            inp.noise_dev = noise.data_ptr()
        if pieces is not None:
            pieces = np.ascontiguousarray(pieces, dtype=NOISE_PIECE_DTYPE)
            assert noise_bank is not None and noise_bank.is_cuda and noise_bank.dtype == torch.float32 and noise_bank.dim() == 1
            inp.noise_bank_dev, inp.noise_bank_len = noise_bank.data_ptr(), noise_bank.numel()
            inp.pieces_host, inp.n_pieces = pieces.ctypes.data, len(pieces)
            keep.append(pieces)
        if rows is None:
            rows = torch.empty(B * 1024 * shifts, 2, dtype=torch.int32)
        if offsets is None:
            offsets = torch.empty(B + 1, dtype=torch.int64)
        call = lambda r: _lib.mfpa_augment_fingerprint_host(self._h, C.byref(inp), B, T, sample_rate, params.ctypes.data_as(C.c_void_p),
                                                            shifts, C.byref(afp), _ptr(r), r.shape[0], _ptr(offsets))
        rc = call(rows)
        if rc == -4:  # MFPA_ECAP: offsets are valid, retry with the exact size
            rows = torch.empty(int(offsets[-1]), 2, dtype=torch.int32)
            rc = call(rows)
        check(rc)
        del keep
        return rows[: int(offsets[-1])].numpy(), offsets.numpy()

    # ---- S5 --------------------------------------------------------------
    def index_load(self, table, counts, hashesperid, hash_lo: int = 0, hashbits: int = 20, maxtimebits: int = 14):
        """Upload (a hash-range shard of) a reference HashTable.  table: uint32 [n_buckets, depth] numpy
        (rows hash_lo .. hash_lo+n_buckets of the full table), counts int32 [n_buckets], hashesperid uint32."""
        import numpy as np

        table = np.ascontiguousarray(table, dtype=np.uint32)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        hpid = np.ascontiguousarray(hashesperid, dtype=np.uint32)
        assert table.ndim == 2 and counts.shape == (table.shape[0],)
        check(_lib.mfpa_index_load(self._h, table.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                   hash_lo, table.shape[0], table.shape[1], hashbits, maxtimebits,
                                   hpid.ctypes.data_as(C.c_void_p), len(hpid)))
        self.n_tracks = len(hpid)
        self.depth = int(table.shape[1])

    def get_hits(self, hashes):
        """hashes: int32 [n,2] cuda -> int32 [nhits,4] cuda (HashTable.get_hits order)."""
        import torch

        hashes = hashes.contiguous()
        n = hashes.shape[0]
        cap = max(1, n * 128)
        hits = torch.empty(cap, 4, dtype=torch.int32, device=hashes.device)
        nh = torch.zeros(1, dtype=torch.int64, device=hashes.device)
        check(_lib.mfpa_get_hits(self._h, _ptr(hashes), n, _ptr(hits), cap, _ptr(nh), _stream()))
        k = int(nh.item())
        if k > cap:
            hits = torch.empty(k, 4, dtype=torch.int32, device=hashes.device)
            check(_lib.mfpa_get_hits(self._h, _ptr(hashes), n, _ptr(hits), k, _ptr(nh), _stream()))
        return hits[:k]

    def match(self, hashes, nh, params: MatchParams | None = None, max_rows: int = 16):
        """Single-shard match_hashes for a batch: hashes int32 [B,cap,2] cuda, nh int32 [B] cuda ->
        (results int32 [B,max_rows,7], nrows int32 [B])."""
        import torch

        params = params or match_defaults()
        B, cap, _ = hashes.shape
        res = torch.zeros(B, max_rows, 7, dtype=torch.int32, device=hashes.device)
        nrows = torch.empty(B, dtype=torch.int32, device=hashes.device)
        check(_lib.mfpa_match(self._h, _ptr(hashes), _ptr(nh), B, cap, C.byref(params), _ptr(res), _ptr(nrows),
                              max_rows, _stream()))
        return res, nrows

    def match_emit(self, hashes, nh, words_cap: int, words=None, nwords=None):
        """Hits of this context's index shard as (track << 15 | time skew + 16384) words: (words int32 [B,words_cap],
        nwords int32 [B])."""
        import torch

        B, cap, _ = hashes.shape
        if words is None:
            words = torch.empty(B, words_cap, dtype=torch.int32, device=hashes.device)
        if nwords is None:
            nwords = torch.empty(B, dtype=torch.int32, device=hashes.device)
        check(_lib.mfpa_match_emit(self._h, _ptr(hashes), _ptr(nh), B, cap, _ptr(words), words_cap, _ptr(nwords), _stream()))
        return words, nwords

    # ---- peer memory (the exchange fused into the sweep)
    def peer_alloc(self, nbytes: int):
        """(device address, 64-byte IPC handle) of a zeroed buffer other ranks of this node can map."""
        p, h = C.c_void_p(), (C.c_char * 64)()
        check(_lib.mfpa_peer_alloc(self._h, nbytes, C.byref(p), h))
        return int(p.value), bytes(h.raw)

    def peer_open(self, handle: bytes) -> int:
        p, h = C.c_void_p(), (C.c_char * 64).from_buffer_copy(handle)
        check(_lib.mfpa_peer_open(self._h, h, C.byref(p)))
        return int(p.value)

    def peer_close(self, address: int):
        check(_lib.mfpa_peer_close(self._h, C.c_void_p(address)))

    def peer_free(self, address: int):
        check(_lib.mfpa_peer_free(self._h, C.c_void_p(address)))

    def match_emit_peer(self, hashes, nh, peers: PeerSet, words_cap: int):
        """match_emit with the words stored straight into the owner ranks' buffers of `peers` (NVLink stores)."""
        B, cap, _ = hashes.shape
        check(_lib.mfpa_match_emit_peer(self._h, _ptr(hashes), _ptr(nh), B, cap, C.byref(peers), words_cap, _stream()))

    def peer_barrier(self, peers: PeerSet, epoch: int):
        check(_lib.mfpa_peer_barrier(self._h, C.byref(peers), epoch & 0xFFFFFFFF, _stream()))

    def match_owner_at(self, words_addr: int, nwords_addr: int, n_shards: int, B: int, words_cap: int, params: MatchParams,
                       max_rows: int = 16):
        """match_owner on a receive buffer given by address ([n_shards,B,words_cap] words, [n_shards,B] counts)."""
        import torch

        dev = torch.device("cuda", self.device)
        res = torch.zeros(B, max_rows, 7, dtype=torch.int32, device=dev)
        nrows = torch.empty(B, dtype=torch.int32, device=dev)
        check(_lib.mfpa_match_owner(self._h, _ptr(words_addr), _ptr(nwords_addr), n_shards, B, words_cap, C.byref(params),
                                    _ptr(res), _ptr(nrows), max_rows, _stream()))
        return res, nrows

    def match_owner(self, words, nwords, params: MatchParams, max_rows: int = 16):
        """words int32 [n_shards,B,words_cap], nwords int32 [n_shards,B] (every shard's hits of the queries this rank
        owns) -> (results int32 [B,max_rows,7], nrows int32 [B])."""
        import torch

        n_shards, B, words_cap = words.shape
        res = torch.zeros(B, max_rows, 7, dtype=torch.int32, device=words.device)
        nrows = torch.empty(B, dtype=torch.int32, device=words.device)
        check(_lib.mfpa_match_owner(self._h, _ptr(words), _ptr(nwords), n_shards, B, words_cap, C.byref(params), _ptr(res),
                                    _ptr(nrows), max_rows, _stream()))
        return res, nrows

    def match_counts(self, hashes, nh):
        import torch

        B, cap, _ = hashes.shape
        width = (self.n_tracks + 1) // 2 if getattr(self, "packed_counts", False) else self.n_tracks
        counts = torch.empty(B, width, dtype=torch.int32, device=hashes.device)
        check(_lib.mfpa_match_counts(self._h, _ptr(hashes), _ptr(nh), B, cap, _ptr(counts), _stream()))
        return counts

    def match_select(self, counts, params: MatchParams):
        import torch

        B = counts.shape[0]
        cand = torch.zeros(B, params.search_depth, 2, dtype=torch.int32, device=counts.device)
        ncand = torch.empty(B, dtype=torch.int32, device=counts.device)
        check(_lib.mfpa_match_select(self._h, _ptr(counts), B, C.byref(params), _ptr(cand), _ptr(ncand), _stream()))
        return cand, ncand

    def match_collect(self, hashes, nh, cand, ncand, params: MatchParams, list_cap: int = 2048):
        import torch

        B, cap, _ = hashes.shape
        lst = torch.empty(B, list_cap, dtype=torch.int32, device=hashes.device)
        nlist = torch.empty(B, dtype=torch.int32, device=hashes.device)
        check(_lib.mfpa_match_collect(self._h, _ptr(hashes), _ptr(nh), B, cap, _ptr(cand), _ptr(ncand), C.byref(params),
                                      _ptr(lst), list_cap, _ptr(nlist), _stream()))
        return lst, nlist

    def match_align(self, lists, nlists, cand, ncand, params: MatchParams, max_rows: int = 16):
        """lists int32 [n_lists,B,list_cap], nlists int32 [n_lists,B]."""
        import torch

        n_lists, B, list_cap = lists.shape
        res = torch.zeros(B, max_rows, 7, dtype=torch.int32, device=lists.device)
        nrows = torch.empty(B, dtype=torch.int32, device=lists.device)
        check(_lib.mfpa_match_align(self._h, _ptr(lists), _ptr(nlists), n_lists, B, list_cap, _ptr(cand), _ptr(ncand),
                                    C.byref(params), _ptr(res), _ptr(nrows), max_rows, _stream()))
        return res, nrows

    # ---- Dejavu ----------------------------------------------------------
    def dejavu_peaks(self, arr, amp_min: float = 50.0, neighborhood: int = 10, cap: int = 4096):
        """arr [B,F,N] float64|float32 cuda -> (mask uint8 [B,F,N], peaks int32 [B,cap,2] (freq,time), npeaks [B])."""
        import torch

        assert arr.is_cuda and arr.dim() == 3 and arr.dtype in (torch.float64, torch.float32)
        arr = arr.contiguous()
        B, F, N = arr.shape
        mask = torch.empty(B, F, N, dtype=torch.uint8, device=arr.device)
        peaks = torch.empty(B, cap, 2, dtype=torch.int32, device=arr.device)
        npk = torch.empty(B, dtype=torch.int32, device=arr.device)
        check(_lib.mfpa_dejavu_peaks(self._h, _ptr(arr), int(arr.dtype == torch.float64), B, F, N, neighborhood,
                                     float(amp_min), _ptr(mask), _ptr(peaks), cap, _ptr(npk), _stream()))
        return mask, peaks, npk

    def dejavu_psd(self, x):
        """x [B,T] f32 cuda -> mlab.specgram PSD / max, float32 [B,257,(T-256)//256] (fingerprint.py:60-68)."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
        B, T = x.shape
        nd = _lib.mfpa_dejavu_num_frames(T)
        psd = torch.empty(B, BINS, nd, dtype=torch.float32, device=x.device)
        check(_lib.mfpa_dejavu_psd(self._h, _ptr(x), B, T, _row_stride(x), _ptr(psd), _stream()))
        return psd

    def dejavu_log(self, psd, square: bool = False):
        """10 ln(max(p, max/1e6)) - mean per item (fingerprint.py:77-79); p = psd**2 when `square`."""
        import torch

        psd = psd.contiguous()
        B, F, N = psd.shape
        arr = torch.empty_like(psd)
        check(_lib.mfpa_dejavu_log(self._h, _ptr(psd), B, F, N, int(square), _ptr(arr), _stream()))
        return arr

    # ---- evaluation ------------------------------------------------------
    def mask_metrics(self, predicted, gt):
        """float32 cuda masks [B,H,W] -> float64 [4] cuda: #gt, sum pred@gt, #pred, sum gt@pred (metrics.py:10-162)."""
        import torch

        predicted, gt = predicted.contiguous(), gt.contiguous()
        assert predicted.is_cuda and gt.is_cuda and predicted.dtype == gt.dtype == torch.float32
        assert predicted.dim() == 3 and predicted.shape == gt.shape
        out = torch.empty(4, dtype=torch.float64, device=predicted.device)
        B, H, W = predicted.shape
        check(_lib.mfpa_mask_metrics(self._h, _ptr(predicted), _ptr(gt), B, H, W, _ptr(out), _stream()))
        return out

    def psnr_stats(self, pred, target):
        """float64 cuda tensors of equal size -> float64 [3] cuda: sum squared error, min(target), max(target)."""
        import torch

        pred, target = pred.contiguous(), target.contiguous()
        assert pred.is_cuda and target.is_cuda and pred.dtype == target.dtype == torch.float64 and pred.numel() == target.numel()
        out = torch.empty(3, dtype=torch.float64, device=pred.device)
        check(_lib.mfpa_psnr_stats(self._h, _ptr(pred), _ptr(target), pred.numel(), _ptr(out), _stream()))
        return out

    # ---- fused -----------------------------------------------------------
    def fingerprint(self, x, shifts: int, params: AfpParams, out=None, nh=None, cap: int | None = None):
        """x [B,T] f32 cuda -> (hashes [B,cap,2] int32, nh [B] int32), rows unique and
        sorted by (time, hash) like wavfile2hashes."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
        B, T = x.shape
        if cap is None:
            cap = HASHES_PER_FRAME * num_frames(T) * shifts
        if out is None:
            out = torch.empty(B, cap, 2, dtype=torch.int32, device=x.device)
        if nh is None:
            nh = torch.empty(B, dtype=torch.int32, device=x.device)
        check(_lib.mfpa_fingerprint(self._h, _ptr(x), B, T, _row_stride(x), shifts, C.byref(params), _ptr(out),
                                    out.shape[1], _ptr(nh), _stream()))
        return out, nh

    def fingerprint_denoised(self, x, shifts: int, params: AfpParams, denoiser):
        """wavfile2hashes with ``denoising=True, denoising_model="unet"`` (peak_extractor.py:263-269):
        STFT -> /max -> UNet -> picker -> landmarks -> hashes -> shift merge, all on the device."""
        B, T = x.shape
        mag, qmax = self.stft_mag(x, shifts)
        denoiser.denoise_items(mag, qmax, T, shifts)
        rec, _ = self.audfprint_peaks(mag, None, T, shifts, params)
        hashes, nh = self.landmark_hashes(rec, params, sorted_rows=True)
        if shifts > 1:
            hashes, nh = self.merge_shifts(hashes, nh, shifts, mag.shape[1])
        return hashes, nh

    def fingerprint_host(self, x, shifts: int, params: AfpParams, rows=None, offsets=None):
        """Host buffers in, host buffers out (CSR).  x: [B,T] float32 (or int16 PCM, scaled by 1/32768 on
        the device) numpy array or CPU torch tensor (pinned memory makes the chunked copies overlap the kernels).
        -> (rows int32 [total,2], offsets int64 [B+1]) as numpy views."""
        import numpy as np
        import torch

        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int16 if x.dtype == np.int16 else np.float32))
        assert not x.is_cuda and x.dtype in (torch.float32, torch.int16) and x.dim() == 2 and x.is_contiguous()
        entry = _lib.mfpa_fingerprint_host_pcm16 if x.dtype == torch.int16 else _lib.mfpa_fingerprint_host
        B, T = x.shape
        if rows is None:
            rows = torch.empty(B * 1024 * shifts, 2, dtype=torch.int32)
        if offsets is None:
            offsets = torch.empty(B + 1, dtype=torch.int64)
        rc = entry(self._h, _ptr(x), B, T, shifts, C.byref(params), _ptr(rows), rows.shape[0], _ptr(offsets))
        if rc == -4:  # MFPA_ECAP: offsets are valid, retry with the exact size
            rows = torch.empty(int(offsets[-1]), 2, dtype=torch.int32)
            rc = entry(self._h, _ptr(x), B, T, shifts, C.byref(params), _ptr(rows), rows.shape[0], _ptr(offsets))
        check(rc)
        return rows[: int(offsets[-1])].numpy(), offsets.numpy()


class DejavuIndex:
    """Device-resident Dejavu fingerprints table (sorted by hash) with the lookup and vote kernels."""

    def __init__(self, ctx: Context, keys, tails, songs, offsets, n_songs: int):
        import numpy as np

        self.ctx = ctx
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        tails = np.ascontiguousarray(tails, dtype=np.uint16)
        songs = np.ascontiguousarray(songs, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        assert keys.shape == tails.shape == songs.shape == offsets.shape and keys.ndim == 1
        self.n, self.n_songs = len(keys), int(n_songs)
        self._h = C.c_void_p()
        check(_lib.mfpa_dejavu_index_create(ctx.handle, keys.ctypes.data_as(C.c_void_p), tails.ctypes.data_as(C.c_void_p),
                                            songs.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), self.n,
                                            self.n_songs, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.mfpa_dejavu_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def return_matches(self, qkeys, qtails, qstart, qoffsets):
        """Distinct query hashes (uint64 keys, uint16 tails) with their offsets (CSR) ->
        (pairs int32 [n,2] cuda (song, offset difference), dedup int32 [n_songs] cuda)."""
        import numpy as np
        import torch

        dev = self.ctx._dev()
        qk = torch.from_numpy(np.ascontiguousarray(qkeys, dtype=np.uint64).view(np.int64)).to(dev)
        qt = torch.from_numpy(np.ascontiguousarray(qtails, dtype=np.uint16).view(np.int16)).to(dev)
        qs = torch.from_numpy(np.ascontiguousarray(qstart, dtype=np.int32)).to(dev)
        qo = torch.from_numpy(np.ascontiguousarray(qoffsets, dtype=np.int32)).to(dev)
        n_hashes = len(qkeys)
        npairs = torch.zeros(1, dtype=torch.int64, device=dev)
        dedup = torch.zeros(max(self.n_songs, 1), dtype=torch.int32, device=dev)
        cap = max(1024, 64 * max(1, len(qoffsets)))
        while True:
            pairs = torch.empty(cap, 2, dtype=torch.int32, device=dev)
            check(_lib.mfpa_dejavu_return_matches(self._h, _ptr(qk), _ptr(qt), _ptr(qs), _ptr(qo), n_hashes, _ptr(pairs), cap,
                                                  _ptr(npairs), _ptr(dedup), _stream()))
            n = int(npairs.item())
            if n <= cap:
                return pairs[:n], dedup[: self.n_songs]
            cap = n

    def align(self, pairs):
        """pairs int32 [n,2] cuda -> (song, offset difference, votes) of the top match (song -1: none)."""
        import torch

        pairs = pairs.contiguous()
        best = torch.empty(3, dtype=torch.int32, device=self.ctx._dev())
        check(_lib.mfpa_dejavu_align(self._h, _ptr(pairs), pairs.shape[0], _ptr(best), _stream()))
        return tuple(int(v) for v in best.cpu().tolist())


def unet_param_blob(state_dict):
    """Flatten a reference ``UNet(1, 1)`` state_dict (training/unet.py:75-95) into the float32 vector
    ``mfpa_unet_load`` expects: state_dict order, ``num_batches_tracked`` skipped."""
    import numpy as np

    parts = [v.detach().cpu().float().numpy().ravel() for k, v in state_dict.items()
             if not k.endswith("num_batches_tracked")]
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


class UNetDenoiser:
    """The optional UNet spectrogram denoiser on one Context (tcgen05 implicit-GEMM convolutions)."""

    def __init__(self, ctx: Context, state_dict=None, max_chunk: int | None = None):
        self.ctx = ctx
        self._u = C.c_void_p()
        check(_lib.mfpa_unet_create(ctx.handle, C.byref(self._u)))
        if max_chunk:
            check(_lib.mfpa_unet_set_max_chunk(self._u, int(max_chunk)))
        if state_dict is not None:
            self.load(state_dict)

    def load(self, state_dict):
        blob = state_dict if hasattr(state_dict, "ctypes") else unet_param_blob(state_dict)
        check(_lib.mfpa_unet_load(self._u, blob.ctypes.data_as(C.c_void_p), blob.size))

    def close(self):
        if getattr(self, "_u", None) and self._u.value:
            _lib.mfpa_unet_destroy(self._u)
            self._u = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, x, div=None):
        """x: [B,H,W] (or [B,1,H,W]) float32 cuda -> same shape; ``unet(x / div[b])``."""
        import torch

        shape = x.shape
        x3 = x.reshape(shape[0], shape[-2], shape[-1]).contiguous()
        B, H, W = x3.shape
        out = torch.empty_like(x3)
        check(_lib.mfpa_unet_forward(self.ctx.handle, self._u, _ptr(x3), H * W, W, 1, _ptr(div), B, H, W, _ptr(out),
                                     H * W, W, 1, _stream()))
        return out.reshape(shape)

    def denoise_items(self, mag, qmax, T: int, shifts: int):
        """denoise_mag for a [B*shifts] item batch: shifted items analyse x[off:] and have one frame
        fewer (peak_extractor.py:411-413), and the network must see exactly their frames, so each shift
        goes through the UNet as its own strided sub-batch."""
        items, n, pitch = mag.shape
        for s in range(shifts):
            frames = num_frames(T - shift_offset(s, shifts))
            sub = mag[s::shifts]
            div = qmax[s::shifts].contiguous()
            check(_lib.mfpa_unet_forward(self.ctx.handle, self._u, _ptr(sub), shifts * n * pitch, 1, pitch, _ptr(div),
                                         sub.shape[0], BINS, frames, _ptr(sub), shifts * n * pitch, 1, pitch, _stream()))
        return mag

    def denoise_mag(self, mag, qmax):
        """In place on the frame-major magnitudes of ``Context.stft_mag``: mag[i] <- unet(mag[i] / qmax[i])
        (peak_extractor.py:263-269); returns mag.  The picker must then run without the STFT's statistics
        (``audfprint_peaks(mag, None, ...)``)."""
        items, n, pitch = mag.shape
        assert pitch == MAG_PITCH
        check(_lib.mfpa_unet_forward(self.ctx.handle, self._u, _ptr(mag), n * pitch, 1, pitch, _ptr(qmax), items, BINS, n,
                                     _ptr(mag), n * pitch, 1, pitch, _stream()))
        return mag


def conv_bf16(ctx: Context, x, w, scale, shift, relu=True, taps=9, out=None, coff=0, bn=0, mt=0, stages=0, halo_wh=0,
              wres=0):
    """x [N,H,W,Cin] bf16 cuda, w [Cout,taps,Cin] bf16 cuda -> [N,H,W,Cout] bf16 (or a channel slice of `out`)."""
    import torch

    N, H, W, cin = x.shape
    cout = w.shape[0]
    if out is None:
        out = torch.empty(N, H, W, cout, dtype=torch.bfloat16, device=x.device)
    check(_lib.mfpa_conv_bf16(ctx.handle, _ptr(x), N, H, W, cin, _ptr(w), cout, taps, _ptr(scale), _ptr(shift), int(relu),
                              _ptr(out), out.shape[-1], coff, bn, mt, stages, halo_wh, wres, _stream()))
    return out
