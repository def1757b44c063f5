"""Drop-in for the reference's `augmentation` package: `AugmentFP` with the same constructor,
`__call__`, `batch_augment` and `augmentation_pipeline` surface (augmentation/__init__.py:16-101).

Host side (this file): the Bernoulli gates and parameter draws of each transform, in the
reference's RNG call order, and noise / impulse-response assembly.
Device side: the whole chain runs in libmfpa (`mfpa_augment`) on the dumped parameters.
"""
from __future__ import annotations

import os
import random
from pkgutil import extend_path

# `augmentation.utils`, `.transform`, `.composition`, `.transformations.*` of a reference checkout later on
# sys.path stay importable next to this replacement of the package's own namespace (AugmentFP)
__path__ = extend_path(__path__, __name__)

from typing import Any, Dict, List, Optional

import numpy as np
import torch

from musicfpaugment_b200 import lib, runtime

from .constants import DEFAULT_PARAMETERS, IMPULSE_RESPONSE_DIR


class ObjectDict(dict):
    """augmentation/utils.py ObjectDict: dict with attribute access."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


class EmptyPathException(Exception):
    pass


def convert_frequencies_to_mels(f: torch.Tensor) -> torch.Tensor:
    return 2595.0 * torch.log10(1.0 + f / 700.0)


def convert_mels_to_frequencies(m: torch.Tensor) -> torch.Tensor:
    return 700.0 * (10 ** (m / 2595.0) - 1.0)


class Audio:
    """Loader with the reference's semantics for in-memory entries
    ({"samples": Tensor[C, n], "sample_rate": int}, augmentation/utils.py:288-383) and for paths
    (decoded with torchaudio when its backend is available)."""

    def __init__(self, sample_rate: int, mono: bool = True):
        self.sample_rate, self.mono = sample_rate, mono

    @staticmethod
    def rms_normalize(samples: torch.Tensor) -> torch.Tensor:
        rms = samples.square().mean(dim=-1, keepdim=True).sqrt()
        return samples / (rms + 1e-8)

    def _meta(self, file):
        if isinstance(file, dict) and "samples" in file:
            return file["samples"].shape[1], file["sample_rate"]
        import torchaudio

        info = torchaudio.info(str(file["audio"] if isinstance(file, dict) else file))
        return info.num_frames, info.sample_rate

    def get_num_samples(self, file) -> int:
        n, sr = self._meta(file)
        return int(np.floor(n * self.sample_rate / sr))

    def __call__(self, file, sample_offset: int = 0, num_samples: Optional[int] = None) -> torch.Tensor:
        total, sr = self._meta(file)
        off = round(sample_offset * sr / self.sample_rate)
        n = total - off if num_samples is None else round(num_samples * sr / self.sample_rate)
        if off + n > total:
            raise ValueError(f"Sample offset {off} -- number of samples {n} -- total number of samples {total}.")
        if isinstance(file, dict) and "samples" in file:
            data = torch.as_tensor(file["samples"], dtype=torch.float32)[:, off: off + n]
        else:
            import torchaudio

            data, _ = torchaudio.load(str(file["audio"] if isinstance(file, dict) else file), frame_offset=off, num_frames=n)
        if self.mono and data.shape[0] > 1:
            data = data.mean(dim=0, keepdim=True)
        if sr != self.sample_rate:
            import torchaudio

            # the reference resamples with librosa (kaiser_best); host-side detail, SURVEY.md §8(f) item 3
            data = torchaudio.functional.resample(data, sr, self.sample_rate)
        if num_samples is not None:
            if data.shape[-1] > num_samples:
                data = data[:, :num_samples]
            elif data.shape[-1] < num_samples:
                data = torch.nn.functional.pad(data, (0, num_samples - data.shape[-1]))
        return data


class _Transform:
    """Parameter-facing surface of BaseWaveformTransform (augmentation/transform.py:24-165)."""

    bit = 0

    def __init__(self, p: float, sample_rate: Optional[int] = None):
        if not 0 <= p <= 1:
            raise ValueError("p must be in [0, 1]")
        self.p, self.sample_rate = p, sample_rate
        self.transform_parameters: Dict[str, Any] = {}
        self.are_parameters_frozen = False

    def freeze_parameters(self, seed: int = 0) -> None:
        """transform.py:158-165 — seeds the RNGs once; parameters are still re-drawn per call (App. B.1)."""
        self.are_parameters_frozen = True
        random.seed(seed)
        torch.manual_seed(seed)

    def unfreeze_parameters(self) -> None:
        self.are_parameters_frozen = False

    def draw(self, batch_size: int, num_samples: int) -> torch.Tensor:
        """Bernoulli gate then the transform's own parameters for the selected sub-batch (transform.py:101-114)."""
        gate = torch.distributions.Bernoulli(torch.tensor(float(self.p))).sample((batch_size,)).to(torch.bool)
        self.transform_parameters = {"should_apply": gate}
        if gate.any():
            self.randomize_parameters(int(gate.sum()), num_samples)
        return gate

    def randomize_parameters(self, n: int, num_samples: int) -> None:
        pass

    @staticmethod
    def _uniform(lo: float, hi: float, n: int) -> torch.Tensor:
        return torch.distributions.Uniform(torch.tensor(lo, dtype=torch.float32), torch.tensor(hi, dtype=torch.float32),
                                           validate_args=True).sample((n,))


class LowPassFilter(_Transform):
    bit = lib.AUG_LPF

    def __init__(self, min_cutoff_freq=150.0, max_cutoff_freq=7500.0, p=0.5, sample_rate=None):
        super().__init__(p, sample_rate)
        if min_cutoff_freq > max_cutoff_freq:
            raise ValueError("min_cutoff_freq must not be greater than max_cutoff_freq")
        self.min_cutoff_freq, self.max_cutoff_freq = min_cutoff_freq, max_cutoff_freq

    def randomize_parameters(self, n, num_samples):
        """Uniform in mel between ceil(mel(min)) and floor(mel(max)) (pass_filters.py:58-82)."""
        lo = torch.ceil(convert_frequencies_to_mels(torch.tensor(self.min_cutoff_freq, dtype=torch.float32)))
        hi = torch.floor(convert_frequencies_to_mels(torch.tensor(self.max_cutoff_freq, dtype=torch.float32)))
        self.transform_parameters["cutoff_freq"] = convert_mels_to_frequencies(self._uniform(float(lo), float(hi), n))


class HighPassFilter(LowPassFilter):
    bit = lib.AUG_HPF1


class ApplyImpulseResponse(_Transform):
    bit = lib.AUG_IR

    def __init__(self, ir_paths, p=0.5, sample_rate=None):
        super().__init__(p, sample_rate)
        self.ir_paths = list(ir_paths)
        if len(self.ir_paths) == 0:
            raise EmptyPathException("There are no supported audio files found.")
        self.audio = Audio(sample_rate=sample_rate, mono=True)
        # every response is decoded once; with a GPU they also live in a zero-padded device matrix, so a
        # batch's [n, 1, Lmax] tensor is an index gather on the device instead of n decodes + one H2D copy
        self.device_bank = torch.cuda.is_available()
        self._decoded: Dict[Any, torch.Tensor] = {}
        self._keep: List[Any] = []
        self._row: Dict[Any, int] = {}
        self._bank_dev: Optional[torch.Tensor] = None

    def _decode(self, path) -> torch.Tensor:
        key = id(path) if isinstance(path, dict) else str(path)
        ir = self._decoded.get(key)
        if ir is None:
            ir = self._decoded[key] = self.audio(path)[0].contiguous()
            self._row[key] = len(self._row)
            if isinstance(path, dict):
                self._keep.append(path)   # id() keys stay valid only while the object lives
            self._bank_dev = None         # rebuilt on next use
        return ir

    def _bank(self) -> torch.Tensor:
        if self._bank_dev is None:
            rows = sorted(self._row, key=self._row.get)
            lmax = max(len(self._decoded[k]) for k in rows)
            m = torch.zeros(len(rows), lmax)
            for r, k in enumerate(rows):
                m[r, : len(self._decoded[k])] = self._decoded[k]
            self._bank_dev = m.cuda()
        return self._bank_dev

    def randomize_parameters(self, n, num_samples):
        """impulse_response.py:57-71 — random.choices, pad_sequence to the longest."""
        paths = random.choices(self.ir_paths, k=n)
        irs = [self._decode(p) for p in paths]
        lmax = max(len(i) for i in irs)
        if self.device_bank:
            idx = torch.tensor([self._row[id(p) if isinstance(p, dict) else str(p)] for p in paths], device="cuda")
            ir = self._bank()[idx, :lmax].unsqueeze(1)
        else:
            ir = torch.zeros(n, 1, lmax)
            for k, i in enumerate(irs):
                ir[k, 0, : len(i)] = i
        self.transform_parameters["ir"] = ir
        self.transform_parameters["ir_lengths"] = [len(i) for i in irs]
        self.transform_parameters["ir_paths"] = paths


class AddBackgroundNoise(_Transform):
    bit = lib.AUG_NOISE

    def __init__(self, background_paths, min_snr_in_db=3.0, max_snr_in_db=30.0, p=0.5, sample_rate=None):
        super().__init__(p, sample_rate)
        self.background_paths = background_paths
        self.background_paths_to_update = dict(background_paths)
        if len(self.background_paths) == 0:
            raise EmptyPathException("There are no supported audio files found.")
        if min_snr_in_db > max_snr_in_db:
            raise ValueError("min_snr_in_db must not be greater than max_snr_in_db")
        self.min_snr_in_db, self.max_snr_in_db = min_snr_in_db, max_snr_in_db
        self.audio = Audio(sample_rate=sample_rate, mono=True)
        # background audio decoded once into a device-resident bank; noise rows are assembled on the GPU
        self.device_bank = torch.cuda.is_available()   # parameter draws alone (no GPU) keep the host assembly
        self._bank_index, self._bank_new, self._bank_keep, self._bank_len, self._bank_dev = {}, [], [], 0, None

    def random_background(self, target_num_samples: int) -> torch.Tensor:
        """background_noise.py:64-141 — pieces from random scenes/files until the length is reached,
        each RMS-normalised, then the concatenation RMS-normalised again."""
        audio, pieces, missing = self.audio, [], target_num_samples
        while missing > 0:
            scene = random.choice(list(self.background_paths_to_update.keys()))
            path = random.choice(self.background_paths_to_update[str(scene)])
            if isinstance(path, (list, tuple)) and len(path) == 2:  # mix-up pair (:80-113)
                n_bg = min(audio.get_num_samples(path[0]), audio.get_num_samples(path[1]))
                if n_bg >= missing:
                    o1 = random.randint(0, n_bg - missing)
                    s1 = audio(path[0], sample_offset=o1, num_samples=missing)
                    o2 = random.randint(0, n_bg - missing)
                    s2 = audio(path[1], sample_offset=o2, num_samples=missing)
                    piece, missing = 0.5 * (s1 + s2), 0
                else:
                    s1 = audio(path[0])
                    piece, missing = 0.5 * (s1 + audio(path[0])), missing - n_bg  # sic: path[0] twice (App. B.5)
            else:
                n_bg = audio.get_num_samples(path)
                if n_bg >= missing:
                    off = random.randint(0, n_bg - missing)
                    piece, missing = audio(path, sample_offset=off, num_samples=missing), 0
                else:
                    piece, missing = audio(path), missing - n_bg
            pieces.append(piece)
        return audio.rms_normalize(torch.cat([audio.rms_normalize(p) for p in pieces], dim=1))

    # ---- device path: the same draws, the arithmetic in mfpa_noise_assemble -------------
    def _bank_entry(self, path):
        """(offset in the bank, length) of a background file decoded once at the target rate; None when the
        file's own rate differs (the reference resamples each excerpt separately, so excerpts of a resampled
        whole would not be the same numbers - such files stay on the host path)."""
        key = id(path) if isinstance(path, dict) else str(path)
        ent = self._bank_index.get(key)
        if ent is None:
            _, sr = self.audio._meta(path)
            if sr != self.audio.sample_rate:
                ent = False
            else:
                wav = self.audio(path)[0].contiguous()
                ent = (self._bank_len, wav.shape[0])
                self._bank_new.append(wav)
                self._bank_len += wav.shape[0]
            self._bank_index[key] = ent
            if isinstance(path, dict):
                self._bank_keep.append(path)   # id() keys stay valid only while the object lives
        return ent or None

    def _bank(self):
        if self._bank_new:
            parts = ([self._bank_dev] if self._bank_dev is not None else []) + [w.cuda() for w in self._bank_new]
            self._bank_dev, self._bank_new = torch.cat(parts), []
        return self._bank_dev

    def _draw_pieces(self, row: int, target_num_samples: int):
        """random_background's control flow (background_noise.py:64-141) with the same RNG calls, producing
        mfpa_noise_piece rows instead of audio; None if a chosen file cannot come from the bank (the draws
        made so far are then replayed on the host path by the caller, which restores the RNG state)."""
        pieces, missing, audio = [], target_num_samples, self.audio
        while missing > 0:
            scene = random.choice(list(self.background_paths_to_update.keys()))
            path = random.choice(self.background_paths_to_update[str(scene)])
            dst = target_num_samples - missing
            if isinstance(path, (list, tuple)) and len(path) == 2:  # mix-up pair (:80-113)
                ea, eb = self._bank_entry(path[0]), self._bank_entry(path[1])
                if ea is None or eb is None:
                    return None
                n_bg = min(audio.get_num_samples(path[0]), audio.get_num_samples(path[1]))
                if n_bg >= missing:
                    o1 = random.randint(0, n_bg - missing)
                    o2 = random.randint(0, n_bg - missing)
                    pieces.append((ea[0] + o1, eb[0] + o2, row, dst, missing))
                    missing = 0
                else:  # sic: path[0] twice, whole file (App. B.5)
                    pieces.append((ea[0], ea[0], row, dst, ea[1]))
                    missing -= n_bg
            else:
                e = self._bank_entry(path)
                if e is None:
                    return None
                n_bg = audio.get_num_samples(path)
                if n_bg >= missing:
                    off = random.randint(0, n_bg - missing)
                    pieces.append((e[0] + off, -1, row, dst, missing))
                    missing = 0
                else:
                    pieces.append((e[0], -1, row, dst, e[1]))
                    missing -= n_bg
        return pieces

    def random_backgrounds_device(self, n: int, target_num_samples: int):
        """n noise rows [n, 1, T] on the GPU, or None when some source file is not at the target rate."""
        state = random.getstate()
        rows = []
        for r in range(n):
            pcs = self._draw_pieces(r, target_num_samples)
            if pcs is None or sum(p[4] for p in pcs) != target_num_samples:   # a short whole-file piece overshoots: host path
                random.setstate(state)
                return None
            rows.extend(pcs)
        arr = np.zeros(len(rows), dtype=lib.NOISE_PIECE_DTYPE)
        for i, (a, b, q, dst, ln) in enumerate(rows):
            arr[i] = (a, b, q, dst, ln, 0)
        out = runtime.get_context().noise_assemble(self._bank(), arr, n, target_num_samples)
        return out.unsqueeze(1)

    def randomize_parameters(self, n, num_samples):
        bg = self.random_backgrounds_device(n, num_samples) if self.device_bank else None
        if bg is None:
            bg = torch.stack([self.random_background(num_samples) for _ in range(n)])
        self.transform_parameters["background"] = bg
        if self.min_snr_in_db == self.max_snr_in_db:
            self.transform_parameters["snr_in_db"] = torch.full((n,), float(self.min_snr_in_db), dtype=torch.float32)
        else:
            self.transform_parameters["snr_in_db"] = self._uniform(float(self.min_snr_in_db), float(self.max_snr_in_db), n)


class Gain(_Transform):
    bit = lib.AUG_GAIN

    def __init__(self, min_gain_in_db=-18.0, max_gain_in_db=6.0, p=0.5, sample_rate=None):
        super().__init__(p, sample_rate)
        if min_gain_in_db >= max_gain_in_db:
            raise ValueError("max_gain_in_db must be higher than min_gain_in_db")
        self.min_gain_in_db, self.max_gain_in_db = min_gain_in_db, max_gain_in_db

    def randomize_parameters(self, n, num_samples):
        db = self._uniform(float(self.min_gain_in_db), float(self.max_gain_in_db), n)
        self.transform_parameters["gain_factors"] = (10 ** (db / 20)).unsqueeze(1).unsqueeze(1)


class Clipping(_Transform):
    bit = lib.AUG_CLIP

    def __init__(self, min_percentile_threshold=0.0, max_percentile_threshold=1.0, p=0.5, sample_rate=None):
        super().__init__(p, sample_rate)
        assert 0 <= min_percentile_threshold and 1 >= max_percentile_threshold
        if min_percentile_threshold >= max_percentile_threshold:
            raise ValueError("max_percentile_threshold must be higher than min_percentile_threshold")
        self.min_percentile_threshold, self.max_percentile_threshold = min_percentile_threshold, max_percentile_threshold

    def randomize_parameters(self, n, num_samples):
        self.transform_parameters["percentile_threshold"] = self._uniform(
            float(self.min_percentile_threshold), float(self.max_percentile_threshold), n).unsqueeze(1)


class PeakNormalization(_Transform):
    bit = lib.AUG_NORM


class Compose:
    """augmentation/composition.py:12-75 for the AugmentFP chain: draws every transform's parameters in
    order, then runs the whole chain in one libmfpa call."""

    def __init__(self, transforms: List[_Transform], shuffle: bool = False, p: float = 1.0):
        if shuffle:
            raise NotImplementedError("the CUDA chain has the fixed AugmentFP order")
        self.transforms, self.p, self.are_parameters_frozen = list(transforms), p, False

    def freeze_parameters(self, seed: int = 0) -> None:
        self.are_parameters_frozen = True
        for t in self.transforms:
            t.freeze_parameters(seed)

    def unfreeze_parameters(self) -> None:
        self.are_parameters_frozen = False
        for t in self.transforms:
            t.unfreeze_parameters()

    def to(self, device):  # the reference moves nn.Modules; here the device is the libmfpa context's GPU
        return self

    def pack(self, batch_size: int, num_samples: int):
        """Draw all parameters and pack them for mfpa_augment -> (params array, ir tensor|None, noise tensor|None)."""
        arr = np.zeros(batch_size, dtype=lib.AUG_DTYPE)
        ir = noise = None
        hp_slots = iter(("fc1_hz", "fc3_hz"))
        hp_bits = iter((lib.AUG_HPF1, lib.AUG_HPF3))
        for t in self.transforms:
            gate = t.draw(batch_size, num_samples).numpy()
            sel = np.nonzero(gate)[0]
            prm = t.transform_parameters
            if isinstance(t, HighPassFilter):
                slot, bit = next(hp_slots), next(hp_bits)
            elif isinstance(t, LowPassFilter):
                slot, bit = "fc2_hz", lib.AUG_LPF
            else:
                slot, bit = None, t.bit
            arr["apply"][sel] |= bit
            if len(sel) == 0:
                continue
            if slot:
                arr[slot][sel] = prm["cutoff_freq"].numpy()
            elif isinstance(t, ApplyImpulseResponse):
                irs = prm["ir"][:, 0, :]
                ir = torch.zeros(batch_size, irs.shape[-1], device=irs.device)   # the device bank leaves it on the GPU
                ir[torch.from_numpy(sel).to(irs.device)] = irs
                arr["ir_len"][sel] = prm["ir"].shape[-1]  # zero-padded to the longest, like pad_sequence (:65-69)
            elif isinstance(t, AddBackgroundNoise):
                bg = prm["background"][:, 0, :]
                noise = torch.zeros(batch_size, num_samples, device=bg.device)   # the device bank leaves it on the GPU
                noise[torch.from_numpy(sel).to(bg.device)] = bg
                arr["snr_db"][sel] = prm["snr_in_db"].numpy()
            elif isinstance(t, Gain):
                arr["gain_factor"][sel] = prm["gain_factors"].reshape(-1).numpy()
            elif isinstance(t, Clipping):
                arr["clip_p"][sel] = prm["percentile_threshold"].reshape(-1).numpy()
        return arr, ir, noise

    def __call__(self, samples: torch.Tensor = None, sample_rate: Optional[int] = None) -> ObjectDict:
        if not isinstance(samples, torch.Tensor) or samples.dim() != 3:
            raise RuntimeError(
                "torch-audiomentations expects three-dimensional input tensors, with"
                " dimension ordering like [batch_size, num_channels, num_samples]. If your"
                " audio is mono, you can use a shape like [batch_size, 1, num_samples].")
        B, C, T = samples.shape
        if C != 1:
            raise NotImplementedError("AugmentFP is used on mono audio; the CUDA chain is mono")
        arr, ir, noise = self.pack(B, T)
        sr = next((t.sample_rate for t in self.transforms if t.sample_rate), sample_rate) or sample_rate
        ctx = runtime.get_context()
        x = samples[:, 0, :].float().contiguous().cuda()
        # a batch clips every selected row to quantiles over the POOLED selected rows (clipping.py:76-90: torch.quantile
        # with a vector q flattens its input); one row = the per-query rule
        ctx.set_option(lib.OPT_CLIP_POOLED, 1 if B > 1 else 0)
        out = ctx.augment(x, arr, ir.cuda() if ir is not None else None, noise.cuda() if noise is not None else None,
                          sample_rate=int(sr))
        out = out if samples.is_cuda else out.cpu()
        return ObjectDict(samples=out.unsqueeze(1), sample_rate=sample_rate)

    forward = __call__


class AugmentFP(object):
    """Music augmentation class for audio fingerprinting (augmentation/__init__.py:16-101)."""

    def __init__(self, background_paths: Dict[str, List[Any]], sample_rate: int,
                 parameters: Dict[str, float] = DEFAULT_PARAMETERS, impulse_response_dir: Any = IMPULSE_RESPONSE_DIR) -> None:
        if isinstance(impulse_response_dir, (str, os.PathLike)):
            ir_paths = [os.path.join(impulse_response_dir, f) for f in os.listdir(impulse_response_dir) if f.endswith(".wav")]
        else:  # in-memory impulse responses ({"samples": Tensor[1, L], "sample_rate": sr}) for IO-less use
            ir_paths = list(impulse_response_dir)
        self.augmentation_pipeline = Compose(transforms=[
            HighPassFilter(p=parameters["proba_cutoff_freq1"], min_cutoff_freq=parameters["min_cutoff_freq1"],
                           max_cutoff_freq=parameters["max_cutoff_freq1"], sample_rate=sample_rate),
            ApplyImpulseResponse(ir_paths, sample_rate=sample_rate, p=parameters["proba_ir_response"]),
            AddBackgroundNoise(background_paths, p=parameters["proba_snr_in_db"], min_snr_in_db=parameters["min_snr_in_db"],
                               max_snr_in_db=parameters["max_snr_in_db"], sample_rate=sample_rate),
            Gain(p=parameters["proba_gain_in_db"], min_gain_in_db=parameters["min_gain_in_db"],
                 max_gain_in_db=parameters["max_gain_in_db"]),
            Clipping(p=parameters["proba_percentile_threshold"], min_percentile_threshold=0,
                     max_percentile_threshold=parameters["max_percentile_threshold"]),
            LowPassFilter(p=parameters["proba_cutoff_freq2"], min_cutoff_freq=parameters["min_cutoff_freq2"],
                          max_cutoff_freq=parameters["max_cutoff_freq2"], sample_rate=sample_rate),
            HighPassFilter(p=parameters["proba_cutoff_freq3"], min_cutoff_freq=parameters["min_cutoff_freq3"],
                           max_cutoff_freq=parameters["max_cutoff_freq3"], sample_rate=sample_rate),
            PeakNormalization(p=1),
        ])

    def __call__(self, waveform: torch.Tensor) -> Any:
        return self.augmentation_pipeline(waveform.unsqueeze(0)).samples.squeeze(0)

    def batch_augment(self, waveforms: torch.Tensor) -> Any:
        return self.augmentation_pipeline(waveforms).samples.squeeze(0)
