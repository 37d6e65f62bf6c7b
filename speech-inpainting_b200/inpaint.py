"""Batched inpainting pipelines = the reference's driver glue replayed on the device.

  * `InformedInpainter` - I_ea/predict.py:85-207 (mask -> z-norm -> CustomModel -> gather -> cos-sim argmax
    -> centroid paste -> extend_mel -> Generator), batched, ragged mask lengths allowed.
  * `BlindInpainter`    - I_da/scripts/inpainting.py:181-259 (mask -> get_feats x2 -> k-means units ->
    optional splice -> trim -> CodeGenerator).
Integer index arithmetic is done on the host with Python ints exactly as the reference does; everything
that touches samples / frames runs in libsib_b200.so.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops
from .ops import SibError


# ----------------------------------------------------------------------------- a1 (host integers)
def iea_mask_indices(start_sec: float, end_sec: float, sr16: int = 16000, sr22: int = 22050):
    """I_ea/predict.py:85-90, 99-100, 133."""
    mask_ms = int((end_sec - start_sec) * 1000)
    mask_len = mask_ms // 20
    start_mask, end_mask = int(start_sec * sr16), int(end_sec * sr16)
    mask_pos = start_mask // 320
    return dict(mask_len=mask_len, mask_pos=mask_pos, zero16=iea_zero_range(mask_pos, mask_len),
                zero22=(start_mask * sr22 // sr16, end_mask * sr22 // sr16))


def iea_zero_range(mask_pos: int, mask_len: int):
    """I_ea/predict.py:133 / I_ea/dataset/dataset.py:82: [pos*320+80, (pos+L)*320+79-80)."""
    return mask_pos * 320 + 80, (mask_pos + mask_len) * 320 + 79 - 80


def extend_mel(spec: torch.Tensor) -> torch.Tensor:
    """I_ea/hifi_gan/inference_modified.py:16-19: [B,80,T] -> [B,80,floor(T*441/256)] on the device."""
    if not spec.is_cuda:
        raise SibError("extend_mel: CUDA tensor required (no CPU fallback)")
    spec = spec.to(torch.float32).contiguous()
    B, Dm, T = spec.shape
    out = torch.empty(B, Dm, ops.extend_mel_len(T), device=spec.device, dtype=torch.float32)
    ops.extend_mel(spec, out, frame_major=False)
    return out


def shard_batch(n_items: int, world_size: int, rank: int):
    """Contiguous utterance shard of rank `rank`: [rank*n/W, (rank+1)*n/W) with the remainder
    spread over the first ranks.  No data-path collective: replicas only."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ida_matched_frames(n_samples: int, n_code: int, n_f0: int, hop: int = 320) -> int:
    """Code frames that survive `match_length` (I_da/src/multiseries.py:5-73; hops 1/320/80 => unit = 320
    samples) and the 1280-sample tail trim (I_da/scripts/inpainting.py:243-256)."""
    n_unit = min(n_samples // hop, n_code, n_f0 // (hop // 80))
    to_remove = (n_unit * hop) % (16 * 80)
    if to_remove % hop != 0:
        raise SibError("to_remove % code_hop_size != 0 (inpainting.py:245 assert)")
    return n_unit - to_remove // hop


def _i32(v, dev):
    return _i32_many([v], dev)[0]


class _PinnedRing:
    """Small ring of pinned int32 staging buffers: index vectors (mask positions / lengths / offsets / zero ranges) go to
    the device with ONE non-blocking copy per call.  A pageable-memory copy would block the host until everything queued
    before it on the stream has finished, i.e. stall the launch loop once per step."""

    def __init__(self, slots: int = 8):
        self.bufs, self.events, self.i = [None] * slots, [None] * slots, 0

    def upload(self, flat: torch.Tensor, dev):
        k = self.i % len(self.bufs)
        self.i += 1
        if self.events[k] is not None:
            self.events[k].synchronize()       # copy issued `slots` uploads ago: long finished
        if self.bufs[k] is None or self.bufs[k].numel() < flat.numel():
            self.bufs[k] = torch.empty(max(1024, flat.numel()), dtype=torch.int32).pin_memory()
        host = self.bufs[k][: flat.numel()]
        host.copy_(flat)
        out = torch.empty(flat.numel(), dtype=torch.int32, device=dev)
        out.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        return out


_rings = {}


def _i32_many(rows, dev):
    """Python int sequences -> int32 device vectors, one pinned, non-blocking upload for all of them."""
    dev = torch.device(dev)
    if dev.type != "cuda":
        raise SibError("index vectors must go to a CUDA device (no CPU fallback)")
    flat = torch.as_tensor([int(v) for r in rows for v in r], dtype=torch.int32)
    ring = _rings.setdefault((dev.index, torch.cuda.current_stream(dev).cuda_stream), _PinnedRing())
    if flat.numel() == 0:
        return [torch.empty(0, dtype=torch.int32, device=dev) for _ in rows]
    d = ring.upload(flat, dev)
    out, o = [], 0
    for r in rows:
        out.append(d[o:o + len(r)])
        o += len(r)
    return out


class InformedInpainter:
    """I_ea informed inpainting for a batch of utterances (predict.py:85-207)."""

    def __init__(self, model, generator, codebook: torch.Tensor):
        """codebook = ApplyKmeans.C, shape [80, K] (I_ea/dataset/km_label.py:10-34, loss_fn.py:10-14)."""
        self.model, self.generator = model, generator
        dev = model._device
        C = codebook.to(dev, torch.float32)
        all_t = C.t().contiguous()                     # all_embeds_t [K, 80]
        self.center = all_t.mean(dim=0).contiguous()   # center_
        self.cc = (all_t - self.center[None, :]).contiguous()  # all_embeds_t_c
        self.device = dev

    def __call__(self, *args, **kwargs):
        with torch.cuda.device(self.device):   # every launch below goes to the CURRENT device: make it the pipeline's
            return self._call(*args, **kwargs)

    def _call(self, wave16, mel, mask_pos, mask_len, apply_mask: bool = True, normalize: bool = True,
              attention_mask=None, return_int16: bool = False, wave22=None, zero22=None, keep_wave: bool = True):
        """wave16 [B,N] float32 (host or device), mel [B,80,T'] (hop-441 log-mel of the masked 22 kHz wave),
        mask_pos/mask_len: per-utterance frame index / length (ints or sequences).
        Instead of a precomputed `mel` (pass None) the 22.05 kHz rendition `wave22` [B,S] and its zero ranges `zero22`
        (list of (lo, hi), `iea_mask_indices(...)["zero22"]`) can be given: the feature front-end of predict.py:99-104
        (zero-mask, normalize * 0.95, get_mel) then runs on the device too.
        Returns namespace(wave [B,1,S], labels [sum L] int64, mel [B,80,T'] inpainted, offsets)."""
        dev = self.device
        B, N = wave16.shape
        if mel is None:
            if wave22 is None:
                raise SibError("either mel or wave22 must be given")
            from .mel import masked_feature_mel
            mel = masked_feature_mel(wave22.to(dev, non_blocking=True), zero22)
        pos = [int(p) for p in (mask_pos if hasattr(mask_pos, "__len__") else [mask_pos] * B)]
        ln = [int(l) for l in (mask_len if hasattr(mask_len, "__len__") else [mask_len] * B)]
        if len(pos) != B or len(ln) != B:
            raise SibError("mask_pos / mask_len must have one entry per utterance")
        x = wave16.to(dev, torch.float32, non_blocking=True).clone() if wave16.is_cuda else wave16.to(dev, torch.float32, non_blocking=True)
        x = x.contiguous()
        off, acc = [], 0
        for l in ln:
            off.append(acc)
            acc += l
        M = acc
        rng = [iea_zero_range(p, l) for p, l in zip(pos, ln)]
        pos_t, len_t, off_t, lo_t, hi_t = _i32_many([pos, ln, off, [r[0] for r in rng], [r[1] for r in rng]], dev)
        if apply_mask:  # predict.py:133
            ops.zero_ranges(x, lo_t, hi_t)
        if normalize:   # predict.py:136-141 (processor: do_normalize=True, eps 1e-7)
            lengths = None if attention_mask is None else attention_mask.sum(-1).to(dev, torch.int32)
            xn = torch.empty_like(x)
            ops.znorm(x, xn, lengths, 1e-7)
            x = xn
        T = self.model.base_model.config.feat_extract_output_length(N)
        for p, l in zip(pos, ln):
            if p < 0 or l < 0 or p + l > T:
                raise SibError(f"mask frames [{p},{p + l}) outside the {T} encoder frames")
        # predict.py:163-168: outputs = model(inputs)[b, pos:pos+L]; the head (LayerNorm + Linear) is row-wise, so it is
        # evaluated on the gathered frames only (bit-identical rows, sum(L) of them instead of B*T)
        values, _ = self.model.forward_frames(x, attention_mask, pos_t, len_t, off_t, M)
        # a private, DENSE [B, 80, T'] copy on the device (the paste below writes into it; a permuted view would keep its
        # strides through .clone())
        mel_dev = mel.to(dev, torch.float32, non_blocking=True)
        mel_dev = mel_dev.clone(memory_format=torch.contiguous_format) if mel_dev.data_ptr() == mel.data_ptr() else mel_dev.contiguous()
        labels = torch.empty(max(M, 1), dtype=torch.int64, device=dev)[:M]
        if M > 0:
            ops.cos_argmax(values, self.cc, labels)                        # loss_fn.py:44-46
            ops.paste_centroids(mel_dev, self.cc, self.center, labels, pos_t, len_t, off_t)  # predict.py:184-187
        Tp = mel_dev.shape[2]
        feats = torch.empty(B, ops.extend_mel_len(Tp), mel_dev.shape[1], device=dev, dtype=torch.float32)
        ops.extend_mel(mel_dev, feats, frame_major=True)                   # predict.py:189
        # predict.py:203.  keep_wave=False (the streaming loop, which only ships the int16 rendition) skips the private
        # copy of the float waveform: res.wave is then the plan's own buffer, valid until the next call
        y = self.generator.forward_frame_major(feats, clone=keep_wave)
        res = SimpleNamespace(wave=y, labels=labels, mel=mel_dev, offsets=off, values=values)
        if return_int16:                                                   # predict.py:204-206
            res.int16 = torch.empty(y.shape, dtype=torch.int16, device=dev)
            ops.pack_int16(y, res.int16)
        return res


    def stream(self, batches, depth: int = 2):
        with torch.cuda.device(self.device):
            yield from self._stream(batches, depth)

    def _stream(self, batches, depth: int = 2):
        """Pipelined serving loop over an iterable of HOST batches (pinned memory recommended) - the micro-batch executor
        BASELINE config 5 needs (1024 x 10 s does not fit one pass): while batch i computes on the current stream, a copy
        stream uploads batch i+1 into one of `depth` device slots and a third one downloads the int16 result of batch i-1.

        Each batch is a dict(wave16=[B,N] f32, mel=[B,80,T'] f32, mask_pos=, mask_len=) (same meaning as `__call__`).
        Yields, in order, namespace(int16 = pinned host [B,1,S] int16, labels = host int64) once that batch's download has
        completed; the host buffers are recycled after `depth` further batches."""
        from collections import deque
        dev = self.device
        compute = torch.cuda.current_stream(dev)
        # separate upload and download streams: in one in-order copy stream the upload of batch i+1 would queue behind the
        # download of batch i, which waits for compute i - no overlap at all
        # device slots, pinned result buffers and the two copy streams are kept on the pipeline object: a serving loop calls
        # stream() again and again, and cudaHostAlloc / cudaMalloc inside it would cost milliseconds per call.  One stream()
        # generator per pipeline object at a time.
        cache = self.__dict__.setdefault("_stream_cache", {})
        if depth not in cache:
            cache[depth] = ([None] * depth, [None] * (depth + 1), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        slots, hosts, copy, down = cache[depth]
        pending = deque()
        for sl in slots:
            if sl is not None:          # whatever ran on the compute stream since the last call is ordered before the uploads
                sl.free = torch.cuda.Event()
                sl.free.record(compute)

        def finish(item):
            fin, res, h_pcm, h_lab = item
            fin.synchronize()
            return SimpleNamespace(int16=h_pcm, labels=h_lab, device_result=res)

        for i, b in enumerate(batches):
            wave, mel = b["wave16"], b["mel"]
            sl = slots[i % depth]
            if sl is None or sl.wave.shape != wave.shape or sl.mel.shape != mel.shape:
                sl = slots[i % depth] = SimpleNamespace(wave=torch.empty(wave.shape, device=dev, dtype=torch.float32),
                                                        mel=torch.empty(mel.shape, device=dev, dtype=torch.float32), free=None)
                # the allocator may hand out memory that earlier work on the compute stream is still using: order the
                # first upload into a fresh slot after everything queued there so far
                sl.free = torch.cuda.Event()
                sl.free.record(compute)
            with torch.cuda.stream(copy):
                if sl.free is not None:
                    copy.wait_event(sl.free)          # the compute that last read this slot has finished
                sl.wave.copy_(wave, non_blocking=True)
                sl.mel.copy_(mel, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy)
            compute.wait_event(ready)
            res = self._call(sl.wave, sl.mel, b["mask_pos"], b["mask_len"], return_int16=True, keep_wave=False)
            done = torch.cuda.Event()
            done.record(compute)
            sl.free = done
            hb = hosts[i % (depth + 1)]
            if hb is None or hb[0].shape != res.int16.shape or hb[1].shape != res.labels.shape:
                hb = hosts[i % (depth + 1)] = (torch.empty(res.int16.shape, dtype=torch.int16).pin_memory(),
                                               torch.empty(res.labels.shape, dtype=torch.int64).pin_memory())
            with torch.cuda.stream(down):
                down.wait_event(done)
                hb[0].copy_(res.int16, non_blocking=True)
                hb[1].copy_(res.labels, non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(down)
            pending.append((fin, res, hb[0], hb[1]))   # `res` keeps the device buffers alive until the download is done
            while len(pending) >= depth:
                yield finish(pending.popleft())
        while pending:
            yield finish(pending.popleft())


class BlindInpainter:
    """I_da inpainting for a batch of equal-length utterances (scripts/inpainting.py:181-259)."""

    def __init__(self, hubert, code_generator, kmeans_centers: torch.Tensor, layer: int = -1, normalize: bool = False,
                 code_hop_size: int = 320, sampling_rate: int = 16000, emb_as_long: bool = True):
        """emb_as_long: the reference hands the speaker d-vector to the generator as `torch.LongTensor(emb)`
        (scripts/inpainting.py:233,239; the training set does the same, src/dataset.py:437), i.e. TRUNCATED TOWARD ZERO
        to integers before it is concatenated - the shipped checkpoints were trained on those values.  True (default)
        reproduces that; False feeds the raw float d-vector (a deliberate divergence from the reference)."""
        self.hubert, self.gen = hubert, code_generator
        self.emb_as_long = emb_as_long
        self.device = hubert._device
        self.mu = kmeans_centers.to(self.device, torch.float32).contiguous()  # [K, H]
        with torch.cuda.device(self.device):
            self._mu_packed, self._mu_bias = ops.kmeans_pack(self.mu)
        self.layer, self.normalize = layer, normalize
        self.hop, self.sr = code_hop_size, sampling_rate

    def get_feats(self, x):
        """HubertFeatureReader.get_feats (hubert_feature_reader.py:44-67), batched over equal lengths."""
        if self.normalize:
            xn = torch.empty_like(x)
            ops.znorm(x, xn, None, 1e-5)  # F.layer_norm(x, x.shape)
            x = xn
        feats = []
        for s in range(0, x.shape[1], 1_600_000):
            f, _ = self.hubert.extract_features(x[:, s:s + 1_600_000].contiguous(), padding_mask=None, mask=False,
                                                output_layer=self.layer)
            feats.append(f)
        return feats[0] if len(feats) == 1 else torch.cat(feats, 1)

    def units(self, feats):
        B, T, H = feats.shape
        labels = torch.empty(B * T, dtype=torch.int64, device=self.device)
        # kmeans_model.predict (inpainting.py:204-205): fp32 GEMM + row argmax (sklearn's float32 formulation)
        ops.kmeans_assign(feats.reshape(B * T, H), self._mu_packed, self._mu_bias, labels)
        return labels.view(B, T)

    def __call__(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return self._call(*args, **kwargs)

    def _call(self, wave, mask_size: int, f0_code=None, emb=None, informed: bool = True, return_int16: bool = False,
              f0=None):
        """`f0` = the continuous f0 track [B, 1, frames at hop 80] the reference passes (inpainting.py:231, quantised
        inside CodeGenerator by the frozen f0 VQ-VAE); `f0_code` = already quantised bins [B, frames / 16] instead."""
        dev = self.device
        if (f0 is None) == (f0_code is None) or emb is None:
            raise SibError("BlindInpainter needs emb and exactly one of f0= (continuous track) / f0_code= (pitch bins)")
        y = wave.to(dev, torch.float32).contiguous()
        B, N = y.shape
        if self.emb_as_long:                                                 # inpainting.py:233 `torch.LongTensor(emb)`
            emb = emb.to(torch.int64)
        frame_start = int(self.sr * 3 / 2)                                  # inpainting.py:187
        y_inp = y.clone()
        ops.zero_ranges(y_inp, _i32([frame_start] * B, dev), _i32([frame_start + mask_size] * B, dev), add_eps=1e-6)  # :188-191
        code = self.units(self.get_feats(y))                                # :195-205
        code_inp = self.units(self.get_feats(y_inp))
        if informed:                                                         # :207-214
            a, b = frame_start // self.hop, (frame_start + mask_size) // self.hop
            code_inp[:, :a] = code[:, :a]
            code_inp[:, b:] = code[:, b:]
        # match_length over (audio hop 1, code hop 320, f0 hop 80) then drop the tail so that the audio is a
        # multiple of 1280 samples (multiseries.py:5-73, inpainting.py:217-256)
        n_f0 = f0.shape[-1] if f0 is not None else 16 * f0_code.shape[1]         # one f0 bin = 16 f0 frames
        n_code = ida_matched_frames(N, code.shape[1], n_f0, self.hop)
        code, code_inp = code[:, :n_code].contiguous(), code_inp[:, :n_code].contiguous()
        if f0 is not None:   # the encoder + quantiser are deterministic: one pass serves both generate() calls
            zp = self.gen.fo_vqvae.encode(f0.to(dev)[..., : n_code * 4]) if getattr(self.gen, "fo_vqvae", None) is not None \
                else None
            if zp is None:
                raise SibError("f0= needs a CodeGenerator configured with an f0_quantizer")
        else:
            zp = f0_code.to(dev)[:, : n_code // 4].contiguous()
        wav_gen = self.gen(code=code, f0_code=zp, emb=emb)                   # :258
        wav_inp = self.gen(code=code_inp, f0_code=zp, emb=emb)               # :259
        res = SimpleNamespace(audio_gen=wav_gen, audio_inp=wav_inp, code=code, code_inpainting=code_inp,
                              audio_mask=y_inp)
        if return_int16:
            res.int16 = torch.empty(wav_inp.shape, dtype=torch.int16, device=dev)
            ops.pack_int16(wav_inp, res.int16)
        return res


def predict_files(pipe: InformedInpainter, wave_paths, start_sec: float, end_sec: float, save_dir=None):
    """`predict(...)` of I_ea/predict.py:66-207 from wave files to wave files, for a batch of utterances:
    `librosa.load(path, sr=22050)` + `librosa.load(path, sr=16000)` (:79-80) through the device resampler, the host
    integer mask arithmetic (:85-90, 99-100), the 22 kHz feature front-end (:99-104), the informed pipeline (:133-203)
    and the int16 conversion (:204-206).  Files are grouped by sample count; every group is one batched pass.
    With `save_dir`, `<save_dir>/<stem>/{masked,inpainted}.wav` are written (:134, :207; 16 kHz / 22.05 kHz PCM-16).

    Returns a list (in the order of `wave_paths`) of namespaces(int16 [S], labels, mask_pos, mask_len)."""
    import os
    from .audio import load_wav_batch, write_wav
    paths = [os.fspath(p) for p in wave_paths]
    idx = iea_mask_indices(start_sec, end_sec)
    both = load_wav_batch(paths, (16000, 22050), pipe.device)
    (w16, n16), (w22, n22) = both[16000], both[22050]
    results = [None] * len(paths)
    shapes = list(zip(n16.tolist(), n22.tolist()))
    for n, s22 in sorted(set(shapes)):
        rows = [i for i, v in enumerate(shapes) if v == (n, s22)]
        x16 = w16[rows, :n].contiguous()
        x22 = w22[rows, :s22].contiguous()
        res = pipe(x16, None, idx["mask_pos"], idx["mask_len"], wave22=x22, zero22=[idx["zero22"]] * len(rows),
                   return_int16=True)
        masked16 = None
        if save_dir is not None:
            lo, hi = idx["zero16"]
            m = x16.clone()
            ops.zero_ranges(m, _i32([lo] * len(rows), pipe.device), _i32([hi] * len(rows), pipe.device))
            masked16 = torch.empty(m.shape, dtype=torch.int16, device=pipe.device)
            ops.pack_int16(m, masked16)
        L = idx["mask_len"]
        pcm = res.int16.reshape(len(rows), -1).cpu()
        for j, i in enumerate(rows):
            results[i] = SimpleNamespace(int16=pcm[j], labels=res.labels[j * L:(j + 1) * L].cpu(), mask_pos=idx["mask_pos"],
                                         mask_len=L)
            if save_dir is not None:
                out = os.path.join(os.fspath(save_dir), os.path.splitext(os.path.basename(paths[i]))[0])
                os.makedirs(out, exist_ok=True)
                write_wav(os.path.join(out, "masked.wav"), masked16[j].cpu(), 16000)
                write_wav(os.path.join(out, "inpainted.wav"), pcm[j], 22050)
    return results
