"""Drop-in for afp/audfprint/peak_extractor.py (reference lines cited per method).

Same class, method names, argument meaning and return types; the arithmetic runs in
libmfpa (sm_100a CUDA) through musicfpaugment_b200.lib.  No CPU fallback: without a GPU
and the built library every analysis call raises.
"""
from __future__ import annotations

import pickle
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from musicfpaugment_b200 import lib, runtime

try:  # the reference imports this constant from testing.parameters (peak_extractor.py:19)
    from testing.parameters import WAVEFORM_SAMPLING_RATE
except Exception:  # reference tree not on sys.path
    WAVEFORM_SAMPLING_RATE = 8000


def landmarks2hashes(landmarks_list) -> np.ndarray:
    """(time, bin1, bin2, dtime) -> (time, 20-bit hash) int32 rows.  peak_extractor.py:40-58."""
    lm = np.asarray(landmarks_list, dtype=np.int64).reshape(-1, 4)
    out = np.zeros((lm.shape[0], 2), dtype=np.int32)
    if lm.shape[0]:
        out[:, 0] = lm[:, 0]
        out[:, 1] = ((lm[:, 1] & 255) << 12) | (((lm[:, 2] - lm[:, 1]) & 63) << 6) | (lm[:, 3] & 63)
    return out


def hashes2landmarks(hashes) -> List[Tuple[int, int, int, int]]:
    """Inverse of landmarks2hashes for the value ranges peaks2landmarks produces."""
    h = np.asarray(hashes, dtype=np.int64).reshape(-1, 2)
    b1 = h[:, 1] >> 12
    db = (h[:, 1] >> 6) & 63
    db = np.where(db >= 32, db - 64, db)
    return [(int(t), int(a), int(a + d), int(dt)) for t, a, d, dt in zip(h[:, 0], b1, db, h[:, 1] & 63)]


def locmax(vec, indices: bool = False):
    """peak_extractor.py:61-73."""
    vec = np.asarray(vec)
    nbr = np.zeros(len(vec) + 1, dtype=bool)
    nbr[0] = True
    nbr[1:-1] = np.greater_equal(vec[1:], vec[:-1])
    maxmask = nbr[:-1] & ~nbr[1:]
    return np.nonzero(maxmask)[0] if indices else maxmask


class Audfprint_peaks(object):
    """peak_extractor.py:76-481."""

    def __init__(self, params: Dict[str, Any], denoising: bool = False, denoising_model=None) -> None:
        self.density = params["density"]
        self.target_sr = params["samplerate"]
        self.n_fft = params["n_fft"]
        self.n_hop = params["n_hop"]
        self.shifts = params["shifts"]
        self.f_sd = params["freq-sd"]
        self.maxpksperframe = params["pks-per-frame"]
        self.maxpairsperpeak = 3
        self.mindt = 2
        self.targetdt = 63
        self.targetdf = 31
        self.denoising = denoising
        self.denoising_model = denoising_model
        self.soundfiledur = 0.0
        if self.denoising:
            assert self.denoising_model in ["demucs", "unet"]
            if self.denoising_model == "demucs":
                raise NotImplementedError(
                    "denoising_model='demucs' (training/model.py:163-326) is outside the B200 hot path "
                    "(SURVEY.md section 8f item 4); only the UNet denoiser is built")
        if self.n_fft != lib.N_FFT or self.n_hop != lib.HOP:
            raise ValueError(f"the CUDA path is built for n_fft={lib.N_FFT}, n_hop={lib.HOP} "
                             f"(testing/parameters.py:17-26), got {self.n_fft}/{self.n_hop}")

    # ---- UNet denoiser (the reference builds it at import time from a hard-coded checkpoint,
    # peak_extractor.py:23-27; here it is built on first use from the same file, from
    # $MFPA_UNET_CHECKPOINT, or from a state_dict handed to set_unet_state_dict)
    UNET_CHECKPOINT = "/workspace/src/training/checkpoints/unet_lr_0.001_BS_128/best_epoch.pt"
    _unet_state_dict = None
    _unet = None

    @classmethod
    def set_unet_state_dict(cls, state_dict) -> None:
        cls._unet_state_dict = state_dict
        if cls._unet is not None:
            cls._unet.close()
            cls._unet = None

    def _denoiser(self):
        cls = Audfprint_peaks
        if cls._unet is None:
            sd = cls._unet_state_dict
            if sd is None:
                import os

                import torch

                path = os.environ.get("MFPA_UNET_CHECKPOINT", cls.UNET_CHECKPOINT)
                sd = torch.load(path, map_location="cpu")["model_state_dict"]  # FileNotFoundError like the reference
            cls._unet = lib.UNetDenoiser(self._ctx(), sd)
        return cls._unet

    # ---- parameters -> C struct
    def _afp(self) -> lib.AfpParams:
        p = lib.AfpParams()
        p.a_dec = runtime.a_dec(self.density, self.n_hop)
        p.f_sd = float(self.f_sd)
        p.maxpks = int(self.maxpksperframe)
        p.mindt, p.targetdt, p.targetdf, p.fanout = self.mindt, self.targetdt, self.targetdf, self.maxpairsperpeak
        return p

    def _ctx(self):
        ctx = runtime.get_context()
        if float(self.f_sd) != getattr(ctx, "_f_sd", 30.0):
            ctx.set_spread_table(runtime.spread_table(256, float(self.f_sd)))
            ctx._f_sd = float(self.f_sd)
        return ctx

    @staticmethod
    def _wave(d):
        import torch

        if isinstance(d, torch.Tensor):
            d = d.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(d, dtype=np.float32).reshape(-1))

    # ---- find_peaks (peak_extractor.py:236-311)
    def find_peaks(self, d):
        """-> (pklist [(col, bin)], peaks_mask float32 [256, N], spec float64 [257, N]);
        ([], array([])) for empty input (:253-254)."""
        import torch

        d = self._wave(d)
        if len(d) == 0:
            return [], np.array([])
        ctx = self._ctx()
        x = torch.from_numpy(d).reshape(1, -1).cuda()
        mag, qmax = ctx.stft_mag(x)
        if self.denoising:  # sgram = unet(sgram / max) (:263-269); spec = the float32 network output
            self._denoiser().denoise_mag(mag, qmax)
            rec, _ = ctx.audfprint_peaks(mag, None, x.shape[1], 1, self._afp())
            spec = mag[0, :, : lib.BINS].t().contiguous().cpu().numpy()
        else:
            if not float(qmax[0]) > 0.0:
                print("find_peaks: Warning: input signal is identically zero.")  # :277-280
            rec, _ = ctx.audfprint_peaks(mag, qmax, x.shape[1], 1, self._afp())
            spec = ctx.spec_from_mag(mag, qmax, x.shape[1])[0].cpu().numpy()
        peaks, npk = ctx.peaks_list(rec)
        mask = ctx.peaks_mask(rec)[0].cpu().numpy()
        pk = peaks[0, : int(npk[0])].cpu().numpy()
        return [(int(c), int(b)) for c, b in pk], mask, spec

    # ---- peaks2landmarks (peak_extractor.py:313-346)
    def peaks2landmarks(self, pklist):
        import torch

        if len(pklist) == 0:
            return []
        pk = np.asarray(pklist, dtype=np.int64).reshape(-1, 2)
        n_frames = int(pk[:, 0].max()) + 1
        rec = np.zeros(n_frames, dtype=np.uint64)
        cnt = np.zeros(n_frames, dtype=np.int64)
        for c, b in pk:  # records: byte 0 = count, bytes 1..5 = bins ascending (pklist is column sorted)
            if cnt[c] >= lib.MAX_PKS:
                raise ValueError("more than 5 peaks in one frame")
            rec[c] |= np.uint64(int(b) & 0xFF) << np.uint64(8 * (cnt[c] + 1))
            cnt[c] += 1
        rec |= cnt.astype(np.uint64)
        ctx = self._ctx()
        recd = torch.from_numpy(rec.view(np.int64)).reshape(1, -1).cuda()
        h, nh = ctx.landmark_hashes(recd, self._afp(), sorted_rows=False)
        return hashes2landmarks(h[0, : int(nh[0])].cpu().numpy())

    # ---- batched entry point (what a B200 user should call instead of looping files)
    def waves2hashes(self, waves, shifts: Optional[int] = None):
        """waves: [B, T] float32 (numpy / torch, host or cuda) -> list of int32 [n_i, 2] arrays, each
        exactly what wavfile2hashes returns for that query."""
        import torch

        shifts = self.shifts if shifts is None else shifts
        shifts = 1 if shifts is None or shifts < 2 else int(shifts)
        ctx = self._ctx()
        if self.denoising:
            w = waves if isinstance(waves, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(waves, dtype=np.float32))
            out, nh = ctx.fingerprint_denoised(w.float().contiguous().cuda(), shifts, self._afp(), self._denoiser())
            out, nh = out.cpu().numpy(), nh.cpu().numpy()
            return [out[i, : nh[i]].copy() for i in range(len(nh))]
        if isinstance(waves, torch.Tensor) and waves.is_cuda:
            out, nh = ctx.fingerprint(waves.float().contiguous(), shifts, self._afp())
            out, nh = out.cpu().numpy(), nh.cpu().numpy()
            return [out[i, : nh[i]].copy() for i in range(len(nh))]
        w = waves.detach().cpu().numpy() if isinstance(waves, torch.Tensor) else np.asarray(waves)
        rows, offs = ctx.fingerprint_host(np.ascontiguousarray(w, dtype=np.float32), shifts, self._afp())
        return [rows[offs[i]: offs[i + 1]].copy() for i in range(len(offs) - 1)]

    # ---- file level (peak_extractor.py:348-481)
    def _read(self, filename: str):
        import torch

        ext = filename.split(".")[-1]
        if ext == "pkl":
            with open(filename, "rb") as f:
                d = torch.tensor(pickle.load(f))
            if WAVEFORM_SAMPLING_RATE != self.target_sr:
                from torchaudio.transforms import Resample

                d = Resample(WAVEFORM_SAMPLING_RATE, self.target_sr)(d)
            return d
        if ext == "mp3":
            import torchaudio  # decoding is host IO, outside the hot path
            from torchaudio.transforms import Resample

            d, sr = torchaudio.load(filename)
            return Resample(sr, self.target_sr)(d.mean(axis=0))
        raise ValueError(f"unsupported file type: {filename}")

    def wavfile2peaks(self, filename: str, shifts: Optional[int] = None, get_masks_waveforms: bool = False):
        d = self._read(filename)
        if d.max().to(float) == 0:
            print("error with filename: ", filename)  # :391-392
        self.soundfiledur = len(d) / self.target_sr
        if shifts is None or shifts < 2:
            peaks, peaks_mask, sgram = self.find_peaks(d)
            if get_masks_waveforms:
                return peaks_mask, d, sgram
            return peaks
        peaklists = []
        for shift in range(shifts):
            shiftsamps = int(shift / self.shifts * self.n_hop)  # :412
            peaklists.append(self.find_peaks(d[shiftsamps:])[0])
        return peaklists

    def wavfile2hashes(self, filename: str):
        d = self._read(filename)
        if d.max().to(float) == 0:
            print("error with filename: ", filename)
        self.soundfiledur = len(d) / self.target_sr
        if len(d) == 0:
            return np.hstack([]).astype(np.int32)  # what the reference does on empty input (:434-435)
        return self.waves2hashes(self._wave(d)[None, :])[0]

    def ingest(self, hashtable, filename: str) -> Tuple[float, int]:
        hashes = self.wavfile2hashes(filename)
        hashtable.store(filename, hashes)
        return self.soundfiledur, len(hashes)
