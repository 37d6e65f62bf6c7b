"""The reference's CPU path assembled from the REAL modules where they can be imported, for bench.py's reference arm.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py): only `bench.py --impl reference`, bench.py's `cpu_baseline` leg and
tests may import this.

  * HuBERT: `transformers.HubertModel` - the reference's own dependency (I_ea/model.py:8,32,40; pinned 4.35.0 in
    requirements.txt:8, whatever version the image carries here) - built from a config and loaded with the seeded state
    dict, followed by `final_layers = Sequential(LayerNorm(H), Linear(H, 80))` exactly as I_ea/model.py:75-89 wires it.
    `transformers` is installed in the image, so this part is the real thing on the GPU box too.
  * HiFi-GAN: `I_ea.hifi_gan.models.Generator` imported from /root/reference through the shims of
    oracle/make_golden.py when that tree exists (the build container); on the GPU box (/root/reference absent)
    the pinned port `oracle.hifigan_ref.generator_forward` (max-abs 5e-7 against that module, tests/golden).
  * glue (mask, z-norm, gather, cos-sim argmax, paste, extend_mel): `oracle.glue_ref`, pinned line by line against
    predict.py / loss_fn.py / inference_modified.py by oracle/make_golden.py.
"""
from __future__ import annotations

import os

import torch


def _hf_config(c):
    from transformers import HubertConfig
    return HubertConfig(hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers,
                        num_attention_heads=c.num_attention_heads, intermediate_size=c.intermediate_size,
                        feat_extract_norm=c.feat_extract_norm, conv_bias=c.conv_bias,
                        do_stable_layer_norm=c.do_stable_layer_norm, conv_dim=list(c.conv_dim),
                        conv_kernel=list(c.conv_kernel), conv_stride=list(c.conv_stride),
                        num_conv_pos_embeddings=c.num_conv_pos_embeddings,
                        num_conv_pos_embedding_groups=c.num_conv_pos_embedding_groups, attn_implementation="eager")


class ReferenceIea:
    """predict.py:85-207 on the CPU with the reference's modules.  `parts` says which pieces are the real thing."""

    def __init__(self, sd, ocfg, gparams, gcfg, C):
        from . import hifigan_ref
        from .params import fold_weight_norm
        self.ocfg, self.gcfg, self.C = ocfg, gcfg, C
        self.parts = {}
        try:
            from transformers import HubertModel
            hub = HubertModel(_hf_config(ocfg)).eval()
            hub.load_state_dict({k[len("base_model."):]: v for k, v in sd.items() if k.startswith("base_model.")}, strict=True)
            head = torch.nn.Sequential(torch.nn.LayerNorm(ocfg.hidden_size), torch.nn.Linear(ocfg.hidden_size, 80)).eval()
            head.load_state_dict({k[len("final_layers."):]: v for k, v in sd.items() if k.startswith("final_layers.")})
            self._model = lambda x: head(hub(x).last_hidden_state)            # I_ea/model.py:80-89
            import transformers
            self.parts["hubert"] = f"transformers.HubertModel {transformers.__version__} (real dependency)"
        except Exception as e:   # pragma: no cover - transformers is part of the image
            from . import hubert_ref
            self._model = lambda x: hubert_ref.custom_model_forward(sd, ocfg, x)
            self.parts["hubert"] = f"oracle port ({type(e).__name__})"
        self._gen = None
        if os.path.isdir(os.environ.get("SIB_REFERENCE", "/root/reference")) and gcfg.model_in_dim == 80:
            try:
                from .make_golden import AttrDict, import_reference
                import_reference()
                from I_ea.hifi_gan.models import Generator
                gen = Generator(AttrDict(gcfg.as_attrdict()))
                gen.load_state_dict({k: v for k, v in gparams.items() if not k.startswith("emb_")}, strict=True)
                gen.eval()
                gen.remove_weight_norm()                                            # predict.py:122
                self._gen = gen
                self.parts["generator"] = "I_ea.hifi_gan.models.Generator imported from /root/reference (real module)"
            except Exception as e:   # pragma: no cover
                self.parts["generator"] = f"oracle port ({type(e).__name__}: reference import failed)"
        if self._gen is None:
            folded = fold_weight_norm(gparams)
            self._gen = lambda feats: hifigan_ref.generator_forward(folded, gcfg, feats)
            self.parts.setdefault("generator", "oracle port of I_ea/hifi_gan/models.py:Generator (reference tree absent on this box)")

    @property
    def kind(self) -> str:
        return "reference" if all("real" in v for v in self.parts.values()) else "port"

    def __call__(self, wave, mel, pos, ln):
        from . import glue_ref
        with torch.no_grad():
            x = wave.clone()
            for b in range(x.shape[0]):
                lo, hi = glue_ref.iea_zero_range_from_frames(pos[b], ln[b])
                x[b, lo:hi] = 0                                                      # predict.py:133
            out = self._model(glue_ref.processor_znorm(x))                           # :136-163
            labels = [glue_ref.cos_sim_argmax(v, self.C) for v in glue_ref.gather_mask_frames(out, pos, ln)]
            feats = glue_ref.extend_mel(glue_ref.paste_centroids(mel, self.C, labels, pos))   # :184-189
            return self._gen(feats), torch.cat(labels) if labels else torch.empty(0, dtype=torch.int64)


class ReferenceIda:
    """I_da/scripts/inpainting.py:181-259 on the CPU: HuBERT features of the clean and the masked signal, k-means units,
    two CodeGenerator passes.  HuBERT = transformers.HubertModel as the stand-in for the (absent, un-vendored) fairseq
    model; generator = pinned port of I_da/src/models.py (the reference CodeGenerator needs a GPU, vq.py:22)."""

    def __init__(self, hp, ocfg, gp, gcfg, mu):
        self.hp, self.ocfg, self.gp, self.gcfg, self.mu = hp, ocfg, gp, gcfg, mu
        self.parts = {}
        try:
            from transformers import HubertModel
            import transformers
            hub = HubertModel(_hf_config(ocfg)).eval()
            hub.load_state_dict(hp, strict=True)
            self._feats = lambda x: hub(x).last_hidden_state
            self.parts["hubert"] = f"transformers.HubertModel {transformers.__version__} (real dependency; fairseq absent)"
        except Exception as e:   # pragma: no cover
            from . import hubert_ref
            self._feats = lambda x: hubert_ref.hubert_forward(hp, ocfg, x)
            self.parts["hubert"] = f"oracle port ({type(e).__name__})"
        self.parts["generator"] = "oracle port of I_da/src/model.py:CodeGenerator (the reference module needs a GPU, vq.py:22)"
        self.kind = "port"

    def __call__(self, wave, mask_size, zp, emb):
        from . import glue_ref, hifigan_ref
        outs = []
        with torch.no_grad():
            for b in range(wave.shape[0]):                      # the reference script handles one utterance per call
                y = wave[b].numpy()
                y_inp, _ = glue_ref.ida_mask(y, mask_size)
                f = self._feats(torch.from_numpy(y)[None])[0]
                f_inp = self._feats(torch.from_numpy(y_inp.astype("float32"))[None])[0]
                code, code_inp = glue_ref.kmeans_predict_f32(f, self.mu), glue_ref.kmeans_predict_f32(f_inp, self.mu)
                n = glue_ref.ida_matched_frames(wave.shape[1], code.shape[0], 4 * code.shape[0])
                e = glue_ref.ida_emb_longtensor(emb[b:b + 1])
                for c in (code, code_inp):                      # inpainting.py:258-259: two generate() calls
                    outs.append(hifigan_ref.code_generator_forward(self.gp, self.gcfg, c[None, :n], zp[b:b + 1, : n // 4], e))
        return outs
