"""CPU oracle for the Speech-Inpainting inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`speech-inpainting_b200/`) imports this directory.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker / the timed CPU baseline.

It is a plain-PyTorch fp32 (CPU) restatement of what the reference computes on
the north-star path (masked 16 kHz waveform -> HuBERT -> head / codebook assign
-> HiFi-GAN -> waveform).  Every function cites the reference file:line it
follows (`HF:` = transformers/models/hubert/modeling_hubert.py, the un-vendored
third-party dependency where HuBERT's arithmetic lives; pinned 4.35.0 in the
reference's requirements.txt:8, 5.5.0 installed here).

Pinning status (see DESIGN.md "Oracle"):
  * mask index arithmetic      - pinned by the reference's own golden pair
                                 I_ea/prediction/LJ050-0271/{orig,masked}.wav
                                 (tests/golden/mask_golden.json).
  * HuBERT forward             - pinned against transformers.HubertModel run in
                                 this container (tests/golden/hubert_*.npz, and
                                 live in tests/test_oracle_pins.py since
                                 transformers is an installed library).
  * HiFi-GAN Generator (I_ea and I_da), extend_mel, cos_sim, CodeGenerator
    front, mel_spectrogram      - pinned against the reference's own Python
                                 modules imported from /root/reference by
                                 oracle/make_golden.py (tests/golden/*.npz).
  * librosa mel filterbank and fairseq-vs-HF equivalence - parity unpinned
    (librosa / fairseq are absent from the image and the reference tree).
"""
