"""Model-level boundary: the ``mdl.name in {'igrnd','vgrnd','vog'}`` x ``conc_type in {'temp','spat'}``
nn.Modules with the reference's constructor protocol ``cls(cfg, comm)`` (code/mdl_base.py:11-75),
batch-dict ``forward(inp) -> {'mdl_outs', 'mdl_outs_eval'}`` (code/mdl_conc_single.py:68-127) and
state_dict key names, so ``code/main_dist.py`` can build, ``.to(device)``,
DDP-wrap, checkpoint and call them unchanged.

What runs where
  language side (a13)   libvog_b200: embedding gather, tcgen05 input/output projections, persistent LSTM
                        recurrence kernel (code/mdl_vog.py:67-140,250-283); the exact 'fp32x' mode runs the same
                        recurrence kernel between fp32 CUDA-core GEMMs (torch + cuDNN only for its SEP variant)
  everything else       libvog_b200 CUDA kernels through vognet_pytorch_b200.ops:
      prop/seg encoders                       code/mdl_vog.py:291-314
      prop|seg concat                         code/mdl_conc_single.py:50-66,156-174
      object transformer + rank-1 rel. bias   code/mdl_vog.py:456-523
      vis|lang concat + per-frame regroup     code/mdl_vog.py:316-344,693-699 (index math, never stored
                                              in the tensor-core modes)
      multimodal transformer                  code/mdl_vog.py:681-744
      lin2 scorer, un-regroup, sigmoid*masks  code/mdl_vog.py:675-677, code/mdl_conc_single.py:118-127

``sep``/``svsq`` concatenation (code/mdl_conc_sep.py:13-217) runs through the same kernels: every (query, video)
pair becomes one single-video pseudo-query (``_sep_flatten``), plus the video-level verb head.
"""
import math
import os

import torch
from torch import nn

from . import ops, packing
from .packing import PackCache
from .transformer_code import COMPUTE_MODES, FactoredTokens, RelBias, RelTransformer, Transformer


def P_of(batch):
    """proposals per query of a batch dict"""
    return int(batch['pad_region_feature'].shape[1])


class _LangEncoder(nn.Module):
    """Parameter layout of the reference LSTMEncoder (utils/mdl_srl_utils.py:72-112)."""

    def __init__(self, vocab_size, embed_dim, hidden, num_layers):
        super().__init__()
        self.padding_idx = vocab_size
        self.embed_tokens = nn.Embedding(vocab_size + 1, embed_dim, self.padding_idx)
        self.lstm = nn.LSTM(input_size=embed_dim, hidden_size=hidden, num_layers=num_layers,
                            dropout=0.1 if num_layers > 1 else 0., bidirectional=True)


def _lin_relu(i, o):
    return nn.Sequential(nn.Linear(i, o), nn.ReLU())


def _scorer(i):
    return nn.Sequential(nn.Linear(i, 256), nn.ReLU(), nn.Linear(256, 1))


class VOGNetB200(nn.Module):
    CONC_TYPE = None       # 'spat' | 'temp' | 'sep'
    MAX_GRAPHS = 8         # captured forwards kept per module (one per batch-shape signature, LRU)
    USE_OBJ_TX = True      # VidGrnd / VOGNet
    USE_MUL_TX = True      # VOGNet

    def __init__(self, cfg, comm):
        super().__init__()
        self.cfg = cfg
        self.comm = comm if comm is not None else {}
        m = cfg.mdl
        self.vocab_size = int(comm['vocab_size'])
        self.num_prop_per_frm = int(comm['num_prop_per_frm'])
        self.num_sampled_frm = int(cfg.ds.num_sampled_frm)
        self.vid_w, self.vid_h = float(cfg.ds.resized_width), float(cfg.ds.resized_height)
        self.srl_arg_len = int(cfg.misc.srl_arg_length)
        pe_, se_, le_ = m.vsrl.prop_encode_size, m.vsrl.seg_encode_size, m.vsrl.lang_encode_size
        self.ps_dim = pe_ + se_
        self.vl_dim = self.ps_dim + le_
        self.lang_dim = le_

        # ---- language model (code/mdl_vog.py:160-193)
        self.lstm_encoder = _LangEncoder(self.vocab_size, m.input_encoding_size, m.rnn.rnn_size,
                                         m.rnn.num_layers)
        self.lstm_out_feat_proj = _lin_relu(m.rnn.rnn_size * 2, le_)
        self.srl_arg_words_out_enc = _lin_relu(le_ * 2, le_)
        self.srl_simple_lin = _lin_relu(le_ * 3, le_)               # unused in temp/spat forward
        # ---- visual model (:195-218, 417-454)
        self.prop_encoder = _lin_relu(m.prop_feat_dim, pe_)
        self.seg_encoder = _lin_relu(m.seg_feat_dim, se_)
        self.seg_verb_classf = _scorer(se_ + le_)                    # SEP only
        if self.USE_OBJ_TX:
            self.obj_txf = self._make_tx(self.ps_dim, m.obj_tx)
            self.pe_obj_sub_enc = _lin_relu(5, m.obj_tx.n_heads)
        # ---- fusion model (:220-237, 547-585)
        self.lin2 = _scorer(self.vl_dim)
        self.lin_tmp = _scorer(self.vl_dim)                          # unused in temp/spat forward
        if self.USE_MUL_TX:
            if not (m.mul_tx.one_frm or m.mul_tx.cross_frm):
                raise AssertionError('mul_tx needs one_frm or cross_frm (code/mdl_vog.py:622)')
            if m.mul_tx.cross_frm:
                raise NotImplementedError('mul_tx.cross_frm is dead code in the reference '
                                          '(KeyError at code/mdl_vog.py:659-663)')
            self.mult_txf = self._make_tx(self.vl_dim, m.mul_tx)
            self.pe_mul_sub_enc = _lin_relu(5, m.mul_tx.n_heads)

        self.compute = 'fp32x'
        self.use_cuda_graph = False
        # training: False runs the training forward without dropout (deterministic; what the gradient-parity tests use)
        self.train_dropout = True

    @staticmethod
    def _make_tx(d, tx_cfg):
        cls = RelTransformer if tx_cfg.use_rel else Transformer
        kw = dict(d_hidden=d // 2, n_layers=tx_cfg.n_layers, n_heads=tx_cfg.n_heads,
                  drop_ratio=tx_cfg.attn_drop, pe=False)
        if tx_cfg.use_rel:
            kw['d_pe'] = 5
        return cls(d, 0, 0, **kw)

    def set_compute(self, mode):
        """'fp32x' exact fp32 CUDA cores | 'tf32' tcgen05 tf32 GEMMs + bf16 attention | 'bf16'."""
        if mode not in COMPUTE_MODES:
            raise ValueError(f'compute must be one of {COMPUTE_MODES}')
        self.compute = mode
        for t in ('obj_txf', 'mult_txf'):
            if hasattr(self, t):
                getattr(self, t).set_compute(mode)
        return self

    # -----------------------------------------------------------------------------------------
    # language side through torch + cuDNN: the reference formulation, kept as an in-repo cross-check
    # (tests/test_gpu_lstm.py) and for the SEP variant of the exact 'fp32x' mode
    # -----------------------------------------------------------------------------------------
    def language_encode(self, inp):
        """-> [B, nsrl, lang_dim] (num_verbs == 1 for temp/spat)."""
        words = inp['srl_arg_words_ind']
        B, nv, nsrl, L = words.shape
        flat = words.reshape(B * nv, nsrl * L)
        wm = inp['srl_arg_word_mask'].reshape(B * nv, -1)
        pad = wm == -1
        toks = torch.gather(flat, 1, wm.masked_fill(pad, 0))        # no in-place edit of inp
        toks = toks.masked_fill(pad, self.vocab_size)
        lens = inp['srl_arg_word_mask_len'].reshape(B * nv)
        lens_cpu = lens.tolist() if lens.is_cuda else lens.tolist()
        toks = toks[:, :max(lens_cpu)]
        emb = self.lstm_encoder.embed_tokens(toks).transpose(0, 1)
        packed = nn.utils.rnn.pack_padded_sequence(emb, lens_cpu, enforce_sorted=False)
        out, (hn, _) = self.lstm_encoder.lstm(packed)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, padding_value=0.)
        if self.CONC_TYPE == 'sep':          # 'final_hidden' of lang_encode (code/mdl_vog.py:265-279)
            self._sep_verb = self.lstm_out_feat_proj(torch.cat([hn[-2], hn[-1]], -1))
        full = self.lstm_out_feat_proj(out.transpose(0, 1))          # [B*nv, T, le]
        cap = inp['srl_arg_words_capture'].reshape(B * nv, nsrl, 2)
        D = full.shape[-1]
        st = torch.gather(full, 1, cap[..., 0].unsqueeze(-1).expand(B * nv, nsrl, D))
        en = torch.gather(full, 1, cap[..., 1].unsqueeze(-1).expand(B * nv, nsrl, D))
        enc = self.srl_arg_words_out_enc(torch.cat([st, en], 2))
        enc = enc * inp['srl_arg_inds_msk'].reshape(B * nv, nsrl, 1).float()
        return enc.view(B, nv * nsrl, D)

    def _packs(self):
        pk = self.__dict__.get('_pack')
        if pk is None:
            pk = self.__dict__['_pack'] = PackCache()
        return pk

    def _lang_weights(self, kind):
        """cached tensor-core operands of the language side: per layer the forward|reverse W_ih
        stacked to [8H, in] (low precision) with b_ih+b_hh, and W_hh stacked to [2,4H,H] fp32."""
        lstm = self.lstm_encoder.lstm
        params = [p for p in lstm.parameters()] + [self.lstm_encoder.embed_tokens.weight]

        def build():
            layers = []
            for l in range(lstm.num_layers):
                g = lambda n: getattr(lstm, f'{n}_l{l}').detach()              # noqa: E731
                gr = lambda n: getattr(lstm, f'{n}_l{l}_reverse').detach()     # noqa: E731
                wih = torch.cat([g('weight_ih'), gr('weight_ih')], 0).float().contiguous()
                bias = torch.cat([g('bias_ih') + g('bias_hh'), gr('bias_ih') + gr('bias_hh')], 0).float().contiguous()
                whh = torch.stack([g('weight_hh'), gr('weight_hh')], 0).float().contiguous()
                layers.append((ops.cast_lp(wih, kind), bias, whh))
            # layer 0 sees only embedding rows: its input projection W_ih.emb[tok] + b is a function of the
            # token id alone, so it is tabulated once per weight version (exact fp32) and the per-step GEMM
            # becomes a row gather
            emb = self.lstm_encoder.embed_tokens.weight.detach().float().contiguous()
            wih0 = torch.cat([lstm.weight_ih_l0.detach(), lstm.weight_ih_l0_reverse.detach()], 0).float().contiguous()
            table = ops.sgemm_nt(emb, wih0, layers[0][1])                            # [V+1, 8H], exact fp32
            return (layers, table)
        return self._packs().get(('lang', kind), params, build)

    def language_encode_tc(self, inp):
        """Language side without host synchronisation (CUDA-graph capturable): embedding gather,
        one tcgen05 GEMM per layer for the input projections of all timesteps and both directions,
        the persistent LSTM recurrence kernel, tcgen05 GEMMs for the two output projections.
        Same arithmetic as ``language_encode`` (packed-sequence semantics from the device lengths)."""
        kind = ops.LP_BF16 if self.compute == 'bf16' else ops.LP_TF32
        words = inp['srl_arg_words_ind']
        B, nv, nsrl, L = words.shape
        Bq = B * nv
        wm = inp['srl_arg_word_mask'].reshape(Bq, -1)
        T = wm.shape[1]
        lens = inp['srl_arg_word_mask_len'].reshape(Bq).contiguous()
        layers, table0 = self._lang_weights(kind)
        # token gather + layer-0 input projection in one kernel: rows of the per-token table W_ih.emb[tok] + b
        # (time-major), so the first recurrence starts without a GEMM in front of it
        gx = ops.lang_embed(words.reshape(Bq, nsrl * L), wm, table0, self.vocab_size, ops.LP_NONE)
        x_lp = None
        for l, (wih_lp, bias, whh) in enumerate(layers):
            if l > 0:
                gx, _ = ops.tc_gemm(x_lp, wih_lp, bias=bias)
            x_lp = ops.lstm_layer_fwd(gx, whh, lens, T, Bq, kind)
        if self.CONC_TYPE == 'sep':
            # 'final_hidden' (code/mdl_vog.py:265-279): last layer's forward state at the last word | backward state
            # at the first word (rows are time-major), through lstm_out_feat_proj
            xl = x_lp.view(T, Bq, -1)
            Hh = xl.shape[-1] // 2
            ar = torch.arange(Bq, device=xl.device)
            last = torch.cat([xl[(lens - 1).clamp(min=0), ar, :Hh], xl[0, :, Hh:]], -1).float()
            self._sep_verb = ops.sgemm_nt(last, self.lstm_out_feat_proj[0].weight, self.lstm_out_feat_proj[0].bias,
                                          relu=True)
        full, _ = ops.tc_gemm(x_lp, self._lp_weight('lstm_proj', self.lstm_out_feat_proj[0].weight, kind),
                              bias=self.lstm_out_feat_proj[0].bias, relu=True)               # [T*Bq, le]
        D = full.shape[-1]
        # first / last word of every SRL argument, concatenated and cast (one kernel)
        cat_lp = ops.lang_gather(full, inp['srl_arg_words_capture'].reshape(Bq, nsrl, 2), T, Bq, kind)
        enc, _ = ops.tc_gemm(cat_lp, self._lp_weight('srl_enc', self.srl_arg_words_out_enc[0].weight, kind),
                             bias=self.srl_arg_words_out_enc[0].bias, relu=True)
        # srl_arg_inds_msk product + the low-precision copy the multimodal transformer consumes (one kernel)
        enc, enc_lp = ops.mask_rows(enc, inp['srl_arg_inds_msk'].reshape(Bq * nsrl), kind)
        self._lang_lp = enc_lp
        return enc.view(B, nv * nsrl, D)

    # -----------------------------------------------------------------------------------------
    def _groups(self, ncmp):
        """(nfrm, nppf') of the per-frame multimodal sequences (code/mdl_conc_single.py:24-28,131-135)."""
        if self.CONC_TYPE == 'spat':
            return self.num_sampled_frm, ncmp * self.num_prop_per_frm
        return ncmp * self.num_sampled_frm, self.num_prop_per_frm      # 'sep' arrives here with ncmp = 1

    def forward(self, inp):
        feat = inp['pad_region_feature']
        if not feat.is_cuda:
            raise RuntimeError('vognet_pytorch_b200 runs on CUDA only (no CPU path); move the batch '
                               'and the module to a B200 device')
        if self.training and feat.shape[0] > 0:
            # the training step (utils/trn_utils.py:497-505): one autograd node whose backward is the hand-written
            # kernel chain of vognet_pytorch_b200.training
            from . import training
            return training.forward_train(self, inp)
        sep = self.CONC_TYPE == 'sep'
        if feat.shape[0] == 0:                       # empty batch: nothing to launch
            nsrl = inp['srl_arg_words_ind'].shape[2]
            if sep:
                ncmp = feat.shape[1]
                z = feat.new_zeros(0, ncmp, nsrl, feat.shape[2])
                return {'mdl_outs': z, 'mdl_outs_eval': z.clone(), 'vidf_outs': feat.new_zeros(0, ncmp),
                        'fin_scores_loss': feat.new_zeros(0, ncmp, nsrl), 'fin_scores': feat.new_zeros(0, ncmp)}
            z = feat.new_zeros(0, 1, nsrl, feat.shape[1])
            return {'mdl_outs': z, 'mdl_outs_eval': z.clone()}
        with torch.no_grad():
            if sep:
                inp, (B, ncmp) = self._sep_flatten(inp)
            out = self._forward_fp32x(inp) if self.compute == 'fp32x' else self._forward_tc(inp)
            if sep:
                nsrl, P1 = out['mdl_outs'].shape[2:]
                out = {'mdl_outs': out['mdl_outs'].view(B, ncmp, nsrl, P1),
                       'mdl_outs_eval': out['mdl_outs_eval'].view(B, ncmp, nsrl, P1),
                       'vidf_outs': out['vidf_outs'].view(B, ncmp),
                       'fin_scores_loss': out['fin_scores_loss'].view(B, ncmp, nsrl),
                       'fin_scores': out['fin_scores'].view(B, ncmp)}
            return out

    # -----------------------------------------------------------------------------------------
    # SEP concatenation (code/mdl_conc_sep.py:13-217)
    # -----------------------------------------------------------------------------------------
    _SEP_LANG_KEYS = ('srl_arg_words_ind', 'srl_arg_word_mask', 'srl_arg_word_mask_len', 'srl_arg_words_capture',
                      'srl_arg_inds_msk', 'verb_ind_in_srl')

    def _sep_flatten(self, inp):
        """ConcSEP scores every (query, video) pair on its own: the object transformer sees the video's nfrm*nppf
        proposals (code/mdl_vog.py:505-516 with B*ncmp sequences), the multimodal transformer [B*ncmp*nfrm,
        nsrl*nppf] (:681-744), the language encoding of slot c goes with video c (mdl_conc_sep.py:165-173).  That is
        the single-video TEMP forward over B*ncmp pseudo-queries, so the batch is re-viewed as such (no copies
        for append_everywhere batches)."""
        feat = inp['pad_region_feature']
        if feat.dim() != 4:
            raise ValueError("conc_type 'sep' expects pad_region_feature [B,ncmp,nfrm*nppf,D] "
                             f'(code/mdl_conc_sep.py:131-160), got {tuple(feat.shape)}')
        B, ncmp = feat.shape[:2]
        if inp['new_srl_idxs'].shape[1] != ncmp:
            raise AssertionError('new_srl_idxs and pad_region_feature disagree on ncmp')
        Bq = B * ncmp
        nv = inp['srl_arg_words_ind'].shape[1]
        if nv not in (1, ncmp):
            raise AssertionError(f'{nv} sentence slots for {ncmp} videos (code/mdl_conc_sep.py:165-173 expands 1 -> ncmp)')
        flat = {}
        for k in self._SEP_LANG_KEYS:
            v = inp[k]
            if v.shape[1] == 1 and ncmp > 1:
                v = v.expand(B, ncmp, *v.shape[2:])
            flat[k] = v.reshape(Bq, 1, *v.shape[2:]) if k != 'verb_ind_in_srl' else v.reshape(Bq).contiguous()
        for k in ('pad_region_feature', 'seg_feature_for_frms', 'pad_proposals'):
            flat[k] = inp[k].reshape(Bq, *inp[k].shape[2:])
        flat['new_srl_idxs'] = inp['new_srl_idxs'].reshape(Bq, 1)
        flat['num_cmp_msk'] = inp['num_cmp_msk'].reshape(Bq, 1)
        return flat, (B, ncmp)

    def _sep_heads(self, out, seg_mean, verb, srl_msk, verb_ind, cmp_msk):
        """Video-level verb score (code/mdl_vog.py:365-398) and the fused per-video score (mdl_conc_sep.py:62-117)."""
        logits = out['mdl_outs']
        Bq, _, nsrl, P1 = logits.shape
        sv = torch.cat([verb, seg_mean], -1).contiguous()
        hv = ops.sgemm_nt(sv, self.seg_verb_classf[0].weight, self.seg_verb_classf[0].bias, relu=True)
        vidf = ops.sgemm_nt(hv, self.seg_verb_classf[2].weight, self.seg_verb_classf[2].bias).view(Bq)
        fin_loss, fin_eval = ops.sep_fin_scores(logits.view(Bq, nsrl, P1), vidf, srl_msk.reshape(Bq, nsrl),
                                                verb_ind.reshape(Bq), cmp_msk.reshape(Bq))
        out = dict(out)
        out.update(vidf_outs=vidf, fin_scores_loss=fin_loss, fin_scores=fin_eval)
        return out

    # -----------------------------------------------------------------------------------------
    # tensor-core path ('tf32' / 'bf16')
    # -----------------------------------------------------------------------------------------
    def _lp_weight(self, name, param, kind):
        return self._packs().get((name, kind), (param,), lambda: ops.cast_lp(param.detach().contiguous(), kind))

    def _weights_sig(self):
        ps = self.__dict__.get('_sig_params')
        if ps is None:
            ps = self.__dict__['_sig_params'] = list(self.parameters())
        return packing.params_signature(ps)

    def _refresh_packs(self):
        """Re-pack every stale low-precision / padded weight copy in place (addresses are stable, so captured
        graphs keep working) -> True if some entry had to be re-allocated (graphs must then be recaptured)."""
        caches = [self._packs()] + [getattr(self, t)._exec._pack for t in ('obj_txf', 'mult_txf') if hasattr(self, t)]
        before = sum(c.relocations for c in caches)
        for c in caches:
            c.refresh()
        return sum(c.relocations for c in caches) != before

    def _forward_tc(self, inp):
        feat, seg, props = inp['pad_region_feature'], inp['seg_feature_for_frms'], inp['pad_proposals']
        assert inp['srl_arg_words_ind'].shape[1] == 1, 'temp/spat concatenation has one verb slot per query'
        ncmp = inp['new_srl_idxs'].shape[1]
        if self.use_cuda_graph:
            return self._forward_tc_graph(inp, ncmp)
        lang = self.language_encode_tc(inp)                           # [B, nsrl, 256] fp32
        x, x_lp = self._visual_tc(feat, seg, props, ncmp, nsrl=lang.shape[1])
        out = self._fusion_tc(x, x_lp, lang, props, inp['srl_arg_inds_msk'], inp['num_cmp_msk'], ncmp)
        if self.CONC_TYPE == 'sep':
            out = self._sep_heads(out, self.__dict__.pop('_sep_seg_mean'), self.__dict__.pop('_sep_verb'),
                                  inp['srl_arg_inds_msk'], inp['verb_ind_in_srl'], inp['num_cmp_msk'])
        return out

    def _visual_tc(self, feat, seg, props, ncmp, nsrl=None):
        """prop/seg encoders -> prop|seg rows -> object transformer.  Independent of the language
        side.  -> x [B*P, 512] fp32 and its low-precision copy."""
        kind = ops.LP_BF16 if self.compute == 'bf16' else ops.LP_TF32
        lp_dtype = torch.bfloat16 if kind == ops.LP_BF16 else torch.float32
        dev = feat.device
        B, P, _ = feat.shape
        nppf = self.num_prop_per_frm
        nvf = seg.shape[1]
        assert nvf * nppf == P, (nvf, nppf, P)
        # both halves of the prop|seg row are written by GEMM epilogues (the seg half with row
        # replication over the nppf proposals of its (frame,vid) slot)
        x = torch.empty(B * P, self.ps_dim, device=dev, dtype=torch.float32)
        x_lp = torch.empty(B * P, self.ps_dim, device=dev, dtype=lp_dtype)
        pe_ = self.prop_encoder[0].out_features
        if kind == ops.LP_BF16:
            # bf16 mode: the 2048-wide region features are read ONCE, as the fp32 A operand of a tf32 MMA (the
            # tensor core drops the low mantissa bits - finer than a bf16 rounding), instead of a cast pass
            # (read fp32 + write bf16) followed by a bf16 GEMM
            ops.tc_gemm(feat.reshape(B * P, -1), self._lp_weight('prop_tf32', self.prop_encoder[0].weight, ops.LP_TF32),
                        bias=self.prop_encoder[0].bias, relu=True, out_f32=x[:, :pe_], out_lp=x_lp[:, :pe_])
        else:
            ops.tc_gemm(ops.cast_lp(feat.reshape(B * P, -1), kind),
                        self._lp_weight('prop', self.prop_encoder[0].weight, kind),
                        bias=self.prop_encoder[0].bias, relu=True, out_f32=x[:, :pe_], out_lp=x_lp[:, :pe_])
        ops.tc_gemm(ops.cast_lp(seg.reshape(B * nvf, -1), kind),
                    self._lp_weight('seg', self.seg_encoder[0].weight, kind),
                    bias=self.seg_encoder[0].bias, relu=True, out_f32=x[:, pe_:], out_lp=x_lp[:, pe_:],
                    rep=nppf)
        if self.CONC_TYPE == 'sep':      # seg_feats.mean(dim=-2) of get_seg_verb_feats_to_process (code/mdl_vog.py:374)
            self._sep_seg_mean = x.view(B, nvf, nppf, self.ps_dim)[:, :, 0, pe_:].mean(1)
        if self.USE_OBJ_TX and self.cfg.mdl.obj_tx.to_use:
            otx = self.cfg.mdl.obj_tx
            if otx.one_frm:
                nfrm_o, nppf_o = self._groups(ncmp)
                Bt_o, N_o, fdiv = B * nfrm_o, nppf_o, float(nfrm_o)
            else:
                Bt_o, N_o, fdiv = B, P, 1.0
            bias = None
            if otx.use_rel:
                # projection + the attention's per-key factors in ONE launch (one kernel less on the chain)
                ex = self.obj_txf._exec
                a, ak = ops.pe_project_expand(props.reshape(B * P, props.shape[-1]), self.pe_obj_sub_enc[0].weight,
                                              self.vid_w, self.vid_h, fdiv, Bt_o, N_o, N_o, 1.0 / math.sqrt(ex.d))
                bias = RelBias(a, self.pe_obj_sub_enc[0].bias, N_o, ak=ak, ak_sig=(Bt_o, N_o, ex.H, ex.d))
            x, x_lp = self.obj_txf._exec.run(x.view(Bt_o, N_o, self.ps_dim), bias, self.compute,
                                             x_lp=x_lp, want_lp=True)
            x = x.reshape(B * P, self.ps_dim)
        if self.USE_MUL_TX and self.cfg.mdl.mul_tx.to_use and self.cfg.mdl.mul_tx.use_rel:
            # the multimodal transformer's bias factors depend on the boxes only: projected here, on the visual
            # branch, so they are off the critical path after the language/visual join
            nfrm_m, nppf2_m = self._groups(ncmp)
            if nsrl is not None:
                # ... and so do the attention's per-key factors (sequence = nsrl x nppf2 tokens of one frame, bias
                # period nppf2): expanded in the same launch, so the attention on the fusion chain starts without its
                # expansion pre-kernel.  _fusion_tc falls back to that pre-kernel if the geometry turns out different.
                exm = self.mult_txf._exec
                sig = (B * nfrm_m, nsrl * nppf2_m, exm.H, exm.d)
                self._a_mul, ak = ops.pe_project_expand(props.reshape(B * P, props.shape[-1]),
                                                        self.pe_mul_sub_enc[0].weight, self.vid_w, self.vid_h,
                                                        float(nfrm_m), sig[0], sig[1], nppf2_m, 1.0 / math.sqrt(exm.d))
                self._ak_mul = (ak, sig)
            else:
                self._a_mul = ops.pe_project(props.reshape(B * P, props.shape[-1]), self.pe_mul_sub_enc[0].weight,
                                             self.vid_w, self.vid_h, float(nfrm_m))
        return x, x_lp

    def _fusion_tc(self, x, x_lp, lang, props, srl_msk, cmp_msk, ncmp):
        """vis|lang tokens regrouped per frame -> multimodal transformer -> lin2 -> masked scores."""
        kind = ops.LP_BF16 if self.compute == 'bf16' else ops.LP_TF32
        B, nsrl = lang.shape[0], lang.shape[1]
        P = x.shape[0] // B
        nppf = self.num_prop_per_frm
        nfrm, nppf2 = self._groups(ncmp)
        lang2 = lang.reshape(B * nsrl, self.lang_dim).contiguous()
        if self.USE_MUL_TX and self.cfg.mdl.mul_tx.to_use:
            # token (b,f,s,p') = [vis[b, f*nppf'+p'] | lang[b,s]] is NEVER written: the first layer of the
            # multimodal transformer projects the two factors separately and reads its residual from them
            mtx = self.cfg.mdl.mul_tx
            bias = None
            if mtx.use_rel:
                a = self.__dict__.pop('_a_mul', None)
                ak, ak_sig = self.__dict__.pop('_ak_mul', (None, None))
                if a is None or a.shape[0] != B * P:
                    a = ops.pe_project(props.reshape(B * P, props.shape[-1]), self.pe_mul_sub_enc[0].weight,
                                       self.vid_w, self.vid_h, float(nfrm))
                    ak, ak_sig = None, None
                bias = RelBias(a, self.pe_mul_sub_enc[0].bias, nppf2, ak=ak, ak_sig=ak_sig)
            lang_lp = self.__dict__.pop('_lang_lp', None)       # produced next to `lang` by language_encode_tc
            if lang_lp is None or lang_lp.shape != lang2.shape:
                lang_lp = ops.cast_lp(lang2, kind)
            ft = FactoredTokens(x.contiguous(), x_lp, lang2, lang_lp, nfrm, nsrl, nppf2)
            xm, xm_lp = self.mult_txf._exec.run_factored(ft, bias, self.compute, need_f32=False)
        else:
            xm, xm_lp = ops.build_xmul(x.contiguous(), lang2, B, nfrm, nsrl, nppf2, kind)
        w1 = self._lp_weight('lin2', self.lin2[0].weight, kind)
        # the fused epilogue needs BN = N (one column tile per 128 rows): worth it once those tiles fill the GPU,
        # small problems keep narrow tiles + the tail kernel
        fuse_min = int(os.environ.get('VOG_FUSED_LIN2_MIN_ROWS', 128 * 64))
        if w1.shape[0] <= 256 and w1.shape[0] % 32 == 0 and xm_lp.shape[0] >= fuse_min:
            # lin2[0] GEMM with lin2[2] + inverse regroup + sigmoid * masks fused into its epilogue
            logits, ev = ops.tc_gemm_lin2(xm_lp.reshape(-1, self.vl_dim), w1, self.lin2[0].bias, self.lin2[2].weight,
                                          self.lin2[2].bias, srl_msk.reshape(B, nsrl), cmp_msk, B, nfrm, nsrl, nppf2,
                                          ncmp, nppf, self.num_sampled_frm, self.CONC_TYPE == 'spat')
        else:
            h, _ = ops.tc_gemm(xm_lp.reshape(-1, self.vl_dim), w1, bias=self.lin2[0].bias, relu=True)
            logits, ev = ops.lin2_tail(h, self.lin2[2].weight, self.lin2[2].bias, srl_msk.reshape(B, nsrl),
                                       cmp_msk, B, nfrm, nsrl, nppf2, ncmp, nppf, self.num_sampled_frm,
                                       self.CONC_TYPE == 'spat')
        return {'mdl_outs': logits, 'mdl_outs_eval': ev}

    # -- CUDA-graph execution: the whole forward is ONE graph per (compute, shapes) signature with
    #    two parallel branches - language side (embedding, LSTM, projections) and visual side
    #    (encoders, object transformer) - joined before the multimodal transformer.  Nothing in it
    #    synchronises with the host; inputs are copied into the graph's static buffers.
    PDL_MAX_ROWS = 4096          # proposals per forward (B * P) up to which the captured forward uses dependent launches
    _GRAPH_KEYS = ('pad_region_feature', 'seg_feature_for_frms', 'pad_proposals', 'srl_arg_inds_msk',
                   'num_cmp_msk', 'srl_arg_words_ind', 'srl_arg_word_mask', 'srl_arg_word_mask_len',
                   'srl_arg_words_capture')

    def _graph_for(self, inp, ncmp):
        """The captured forward for this (compute mode, shapes) signature; captured on first use."""
        feat = inp['pad_region_feature']
        keys = self._GRAPH_KEYS + (('verb_ind_in_srl',) if self.CONC_TYPE == 'sep' else ())
        key = (self.compute, ncmp, feat.device.index) + tuple(tuple(inp[k].shape) for k in keys)
        graphs = self.__dict__.setdefault('_graphs', {})
        g = graphs.get(key)
        if g is None:
            # the captured forward's inputs live in ONE flat buffer (256-byte aligned slices, runtime.PackedLayout): a
            # batch that arrives packed the same way is staged with a single copy
            from .runtime import PackedLayout
            lay = PackedLayout({k: inp[k] for k in keys}, first=keys)
            flat = torch.empty(lay.nbytes, dtype=torch.uint8, device=feat.device)
            st = lay.views(flat)
            for k in keys:
                st[k].copy_(inp[k])

            # The language recurrence and the visual branch run as two parallel branches of the graph.  (An SM partition
            # between them - recurrence on a few SMs, persistent GEMMs on the rest - was measured at spat/p100 in round 1,
            # profiles/r1/lstm_trace.txt: 2.47-3.17 ms per step against 1.94 ms unpartitioned, and removed.)
            from . import _lib

            def body(side):
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    lang = self.language_encode_tc(st)
                x, x_lp = self._visual_tc(st['pad_region_feature'], st['seg_feature_for_frms'],
                                          st['pad_proposals'], ncmp,
                                          nsrl=st['srl_arg_words_ind'].shape[1] * st['srl_arg_words_ind'].shape[2])
                cur.wait_stream(side)
                out = self._fusion_tc(x, x_lp, lang, st['pad_proposals'], st['srl_arg_inds_msk'],
                                      st['num_cmp_msk'], ncmp)
                if self.CONC_TYPE == 'sep':
                    out = self._sep_heads(out, self.__dict__.pop('_sep_seg_mean'), self.__dict__.pop('_sep_verb'),
                                          st['srl_arg_inds_msk'], st['verb_ind_in_srl'], st['num_cmp_msk'])
                return out
            # Programmatic dependent launch for the SMALL configurations: their forward is a chain of ~40 short
            # dependent kernels, and letting each start (barrier / tensor-memory set-up, weight loads) while its
            # predecessor drains is worth 1.8 % at spat/gt5 (profiles/r2/pdl_ab.txt); at spat/p100 it measured 0.7 %
            # SLOWER (early CTAs of the next kernel take SMs from a busy one).  VOG_PDL=0 / 1 forces it off / on.
            pdl_env = os.environ.get('VOG_PDL')
            use_pdl = pdl_env == '1' or (pdl_env is None and feat.shape[0] * feat.shape[1] <= self.PDL_MAX_ROWS)
            _lib.lib().vog_debug_pdl(1 if use_pdl else 0)        # thread-local switch of the library
            try:
                side = torch.cuda.Stream(device=feat.device)
                body(side)                              # eager warm-up: packs weights, sets kernel attributes
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream(device=feat.device)
                n0 = _lib.lib().vog_launch_count()
                with torch.cuda.graph(graph, stream=cap):
                    out = body(side)
            finally:
                _lib.lib().vog_debug_pdl(1 if pdl_env == '1' else 0)
            # kernels of libvog_b200 captured into the graph = launches per replay
            g = dict(st=st, graph=graph, out=out, launches=_lib.lib().vog_launch_count() - n0,
                     wsig=self._weights_sig(), flat=flat, prefix=lay.prefix(keys), keys=keys)
            graphs[key] = g
            while len(graphs) > self.MAX_GRAPHS:            # least recently used first (dicts keep insertion order)
                graphs.pop(next(iter(graphs)))
        else:
            graphs[key] = graphs.pop(key)                    # mark as most recently used
        return g

    def graph_input_buffers(self, inp):
        """The static input tensors of the captured forward for batches shaped like ``inp`` (captured now if
        needed).  A caller that writes its batches straight into these tensors (e.g. as the destination of its
        host-to-device copies) and passes the same dict to ``forward`` skips the staging copies: ``forward``
        only copies inputs whose storage differs from the graph's."""
        if self.CONC_TYPE == 'sep':                  # forward() replays the graph on the flattened pseudo-queries
            inp, _ = self._sep_flatten(inp)
        g = self._graph_for(inp, inp['new_srl_idxs'].shape[1])
        buf = dict(g['st'])
        for k, v in inp.items():                    # keys the graph does not read (only shapes matter) pass through
            buf.setdefault(k, v)
        return buf

    def _forward_tc_graph(self, inp, ncmp):
        g = self._graph_for(inp, ncmp)
        # The captured kernels read the PACKED weight copies through baked-in addresses.  When a parameter changed
        # since the last replay (optimizer step, load_state_dict, in-place edit) the copies are re-packed in place;
        # only if one had to move (shape / dtype change) are the graphs dropped and this one captured again.
        wsig = self._weights_sig()
        if g['wsig'] != wsig:
            # wsig[2] folds the parameters' own addresses: biases, LayerNorm weights, W_hh ... are read by the captured
            # kernels straight from the parameter storage, so a parameter that moved (FlatAdam re-homing, .to(),
            # p.data = ...) invalidates the capture itself
            moved = g['wsig'][2] != wsig[2]
            if self._refresh_packs() or moved:
                self.__dict__['_graphs'] = {}
                g = self._graph_for(inp, ncmp)
            g['wsig'] = wsig
        self.graph_launches = g['launches']
        # staging: every input whose storage is not the graph's own buffer is copied in - all of them in one
        # multi-tensor launch per dtype instead of one copy kernel per tensor (nine per forward)
        lay = getattr(inp, 'layout', None)
        if lay is not None and lay.prefix(g['keys']) == g['prefix'] and inp.flat.device == g['flat'].device:
            g['flat'].copy_(inp.flat[:g['flat'].numel()], non_blocking=True)       # packed batch: one copy
            g['graph'].replay()
            return {k: v.clone() for k, v in g['out'].items()}
        dsts, srcs = [], []
        for k, buf in g['st'].items():
            src = inp[k]
            if src.data_ptr() != buf.data_ptr() or src.stride() != buf.stride():
                if src.dtype == buf.dtype and src.device == buf.device and src.is_contiguous():
                    dsts.append(buf), srcs.append(src.view(buf.shape))
                else:
                    buf.copy_(src, non_blocking=True)
        if dsts:
            torch._foreach_copy_(dsts, srcs, non_blocking=True)
        g['graph'].replay()
        return {k: v.clone() for k, v in g['out'].items()}

    def _mask_outputs(self, logits, inp, B, nsrl, ncmp, nppf, P):
        cm = inp['num_cmp_msk'].float()
        if self.CONC_TYPE == 'spat':
            cmsk = cm.view(B, 1, 1, 1, ncmp, 1).expand(B, 1, nsrl, self.num_sampled_frm, ncmp, nppf)
        else:
            cmsk = cm.view(B, 1, 1, ncmp, 1).expand(B, 1, nsrl, ncmp, self.num_sampled_frm * nppf)
        smsk = inp['srl_arg_inds_msk'].float().view(B, 1, nsrl, 1)
        ev = torch.sigmoid(logits) * smsk * cmsk.reshape(B, 1, nsrl, P)
        return {'mdl_outs': logits, 'mdl_outs_eval': ev}

    def _forward_fp32x(self, inp):
        feat, seg, props = inp['pad_region_feature'], inp['seg_feature_for_frms'], inp['pad_proposals']
        B, P, _ = feat.shape
        ncmp = inp['new_srl_idxs'].shape[1]
        nppf = self.num_prop_per_frm
        nvf = seg.shape[1]
        assert nvf * nppf == P, (nvf, nppf, P)
        nv = inp['srl_arg_words_ind'].shape[1]
        assert nv == 1, 'temp/spat concatenation has one verb slot per query'
        nsrl = inp['srl_arg_words_ind'].shape[2]
        if self.CONC_TYPE == 'sep':
            lang = self.language_encode(inp)                          # needs the final LSTM states (verb head): nn.LSTM
        else:
            # the repo's own kernels in exact fp32 (the language half of the exact training forward, dropout off):
            # embedding gather, vog_sgemm projections, the fp32 recurrence kernel - no library call on this path
            from . import training
            lang = training.lang_forward(self, inp, training.Tape(), training.DropCtx(self, False)).view(B, nv * nsrl, -1)

        # prop|seg features: [B*P, 512], prop half written in place by the GEMM
        x = torch.empty(B * P, self.ps_dim, device=feat.device, dtype=torch.float32)
        pe_ = self.prop_encoder[0].out_features
        ops.sgemm_nt(feat.reshape(B * P, -1), self.prop_encoder[0].weight, self.prop_encoder[0].bias,
                     relu=True, out=x[:, :pe_])
        segf = ops.sgemm_nt(seg.reshape(B * nvf, -1), self.seg_encoder[0].weight,
                            self.seg_encoder[0].bias, relu=True)
        x.view(B * nvf, nppf, self.ps_dim)[:, :, pe_:] = segf.unsqueeze(1)
        props2 = props.reshape(B * P, props.shape[-1])

        if self.USE_OBJ_TX and self.cfg.mdl.obj_tx.to_use:
            otx = self.cfg.mdl.obj_tx
            if otx.one_frm:                                           # code/mdl_vog.py:496-504
                nfrm_o, nppf_o = self._groups(ncmp)
                Bt_o, N_o, fdiv = B * nfrm_o, nppf_o, float(nfrm_o)
            else:
                Bt_o, N_o, fdiv = B, P, 1.0                           # code/mdl_vog.py:505-510
            bias = None
            if otx.use_rel:
                a = ops.pe_project(props2, self.pe_obj_sub_enc[0].weight, self.vid_w, self.vid_h, fdiv)
                bias = RelBias(a, self.pe_obj_sub_enc[0].bias, N_o)
                x = self.obj_txf(x.view(Bt_o, N_o, self.ps_dim), bias)
            else:
                x = self.obj_txf(x.view(Bt_o, N_o, self.ps_dim))
            x = x.reshape(B * P, self.ps_dim)

        nfrm, nppf2 = self._groups(ncmp)
        # token (b,f,s,p') = [vis[b, f*nppf'+p'] | lang[b,s]]  (code/mdl_vog.py:316-344,693-699)
        vis = x.view(B, nfrm, 1, nppf2, self.ps_dim).expand(B, nfrm, nsrl, nppf2, self.ps_dim)
        lng = lang.view(B, 1, nsrl, 1, self.lang_dim).expand(B, nfrm, nsrl, nppf2, self.lang_dim)
        xm = torch.cat([vis, lng], -1).view(B * nfrm, nsrl * nppf2, self.vl_dim)
        if self.USE_MUL_TX and self.cfg.mdl.mul_tx.to_use:
            mtx = self.cfg.mdl.mul_tx
            if mtx.use_rel:
                a = ops.pe_project(props2, self.pe_mul_sub_enc[0].weight, self.vid_w, self.vid_h,
                                   float(nfrm))                       # code/mdl_vog.py:710-713
                xm = self.mult_txf(xm, RelBias(a, self.pe_mul_sub_enc[0].bias, nppf2))
            else:
                xm = self.mult_txf(xm)
        h = ops.sgemm_nt(xm.reshape(-1, self.vl_dim), self.lin2[0].weight, self.lin2[0].bias, relu=True)
        lg = ops.sgemm_nt(h, self.lin2[2].weight, self.lin2[2].bias)  # [B*nfrm*nsrl*nppf', 1]
        logits = lg.view(B, nfrm, nsrl, nppf2).transpose(1, 2).reshape(B, 1, nsrl, P)

        # masks (code/mdl_conc_single.py:39-48,118-122,144-154)
        cm = inp['num_cmp_msk'].float()
        if self.CONC_TYPE == 'spat':
            cmsk = cm.view(B, 1, 1, 1, ncmp, 1).expand(B, 1, nsrl, self.num_sampled_frm, ncmp, nppf)
        else:
            cmsk = cm.view(B, 1, 1, ncmp, 1).expand(B, 1, nsrl, ncmp, self.num_sampled_frm * nppf)
        smsk = inp['srl_arg_inds_msk'].float().view(B, 1, nsrl, 1)
        ev = torch.sigmoid(logits) * smsk * cmsk.reshape(B, 1, nsrl, P)
        out = {'mdl_outs': logits, 'mdl_outs_eval': ev}
        if self.CONC_TYPE == 'sep':
            out = self._sep_heads(out, segf.view(B, nvf, -1).mean(1), self.__dict__.pop('_sep_verb'),
                                  inp['srl_arg_inds_msk'], inp['verb_ind_in_srl'], inp['num_cmp_msk'])
        return out


def _variant(name, conc, obj, mul):
    return type(name, (VOGNetB200,), {'CONC_TYPE': conc, 'USE_OBJ_TX': obj, 'USE_MUL_TX': mul,
                                      '__doc__': f'{name}: code/mdl_vog.py:393-398,526-535,747-756'})


ImgGrnd_TEMP = _variant('ImgGrnd_TEMP', 'temp', False, False)
ImgGrnd_SPAT = _variant('ImgGrnd_SPAT', 'spat', False, False)
VidGrnd_TEMP = _variant('VidGrnd_TEMP', 'temp', True, False)
VidGrnd_SPAT = _variant('VidGrnd_SPAT', 'spat', True, False)
VOG_TEMP = _variant('VOG_TEMP', 'temp', True, True)
VOG_SPAT = _variant('VOG_SPAT', 'spat', True, True)
ImgGrnd_SEP = _variant('ImgGrnd_SEP', 'sep', False, False)
VidGrnd_SEP = _variant('VidGrnd_SEP', 'sep', True, False)
VOG_SEP = _variant('VOG_SEP', 'sep', True, True)
