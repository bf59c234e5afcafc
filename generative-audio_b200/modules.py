"""Full-band modules of FullSubNet+ (TSSE attention a3, TCN sequence model a4) and the parameter container of
the sub-band LSTM (a7), with the reference's parameter names so state_dicts are interchangeable.

a3 (TSSE) runs on three fused kernels; a4 (TCN) has two paths: lstm_impl="tc" runs the whole stack channel-last with every
1x1 convolution and the output Linear on the in-house tcgen05 GEMM (forward_tc), lstm_impl="f32" keeps fp32 activations with
the 1x1 convolutions on the library and everything between them on three fused kernels.  The LSTM is never run through
nn.LSTM on the inference path: its weights feed the hand-written kernels via ops.LstmPlan."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

TCN_DILATIONS = (1, 2, 5, 9, 1, 2, 5, 9)  # sequence_model.py:48-57

# attributes that hold DERIVED caches (a ctypes LSTM plan handle, packed / folded / split 16-bit weights): rebuilt on demand from
# the fp32 master parameters, so copy.deepcopy(model), pickling and torch.save(model) drop them instead of choking on a ctypes
# pointer or carrying stale device buffers into the copy
_DERIVED = ("_plan", "_plan_key", "_tplan", "_tplan_key", "_fold", "_fold_key", "_fold16", "_fold16_key", "_tc", "_tc_key")


class _NoDerivedState:
    def __getstate__(self):
        st = dict(self.__dict__)
        for k in _DERIVED:
            if k in st:
                st[k] = None
        return st

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            setattr_ = object.__setattr__
            setattr_(new, k, copy.deepcopy(v, memo))
        return new


def _conv1x1_f32(w, x):
    """1x1 Conv1d of the fp32 path as a TRUE fp32 GEMM: w [O,C,1], x [B,C,T] -> [B,O,T].  F.conv1d goes through cuDNN, whose
    default (torch.backends.cudnn.allow_tf32 = True) is a TF32 tensor-core GEMM with a 10-bit mantissa — measured 4e-3 on
    w_mat for speech input, 40x the fp32 budget; matmul with TF32 forced off is the library's fp32 SGEMM."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return torch.matmul(w[:, :, 0], x)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


class _ConvPoolRelu(nn.Sequential):
    """Keeps the reference's `smallConv1d.0.weight` key layout (attention_model.py:57-71)."""

    def __init__(self, channels, k):
        super().__init__(nn.Conv1d(channels, channels, kernel_size=k, groups=channels))


class ChannelTimeSenseSELayer(nn.Module):
    """attention_model.py:43-98."""

    def __init__(self, num_channels, reduction_ratio=2, kersize=(3, 5, 10)):
        super().__init__()
        self.smallConv1d = _ConvPoolRelu(num_channels, kersize[0])
        self.middleConv1d = _ConvPoolRelu(num_channels, kersize[1])
        self.largeConv1d = _ConvPoolRelu(num_channels, kersize[2])
        self.feature_concate_fc = nn.Linear(3, 1, bias=True)
        self.fc1 = nn.Linear(num_channels, num_channels // reduction_ratio, bias=True)
        self.fc2 = nn.Linear(num_channels // reduction_ratio, num_channels, bias=True)

    def forward(self, x):  # [B,C,T] -> three fused kernels (squeeze via windowed sums, excitation MLP, scale)
        convs = (self.smallConv1d[0], self.middleConv1d[0], self.largeConv1d[0])
        return ops.tsse(x.contiguous(), [c.kernel_size[0] for c in convs], [c.weight for c in convs],
                        [c.bias for c in convs], self.feature_concate_fc.weight, self.feature_concate_fc.bias,
                        self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class TCNBlock(_NoDerivedState, nn.Module):
    """causal_conv.py:67-108 (use_skip_connection=True, causal=False)."""

    def __init__(self, in_channels=257, hidden_channel=512, out_channels=257, kernel_size=3, dilation=1):
        super().__init__()
        self.conv1x1 = nn.Conv1d(in_channels, hidden_channel, 1)
        self.prelu1 = nn.PReLU()
        self.norm1 = nn.GroupNorm(1, hidden_channel, eps=1e-8)
        self.depthwise_conv = nn.Conv1d(hidden_channel, hidden_channel, kernel_size=kernel_size, stride=1,
                                        groups=hidden_channel, padding=(dilation * (kernel_size - 1)) // 2,
                                        dilation=dilation)
        self.prelu2 = nn.PReLU()
        self.norm2 = nn.GroupNorm(1, hidden_channel, eps=1e-8)
        self.sconv = nn.Conv1d(hidden_channel, out_channels, 1)

        self.dilation = dilation
        self._fold = None
        self._fold_key = None

    def _folded(self):
        """GroupNorm2's affine folded into the second 1x1 conv: W2' = W2 diag(gamma2), u = W2 gamma2, vb = W2 beta2 + b2.
        Derived cache, rebuilt when the parameters change; never serialised."""
        ps = (self.sconv.weight, self.sconv.bias, self.norm2.weight, self.norm2.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._fold is None or key != self._fold_key:
            with torch.no_grad():
                W2 = self.sconv.weight[:, :, 0].double()
                w2f = (W2 * self.norm2.weight.double()[None, :]).float()[:, :, None].contiguous()
                u = (W2 @ self.norm2.weight.double()).float().contiguous()
                vb = (W2 @ self.norm2.bias.double() + self.sconv.bias.double()).float().contiguous()
            self._fold, self._fold_key = (w2f, u, vb), key
        return self._fold

    def forward(self, x):
        """x [B,C,T'] -> x + sconv(norm2(prelu2(dwconv(norm1(prelu1(conv1x1(x)))))))  (causal_conv.py:96-108).
        1x1 convolutions: library GEMM; everything between them: 3 fused kernels (prelu_stats, tcn_mid, tcn_out)."""
        x = x.contiguous()
        y1 = _conv1x1_f32(self.conv1x1.weight, x)   # bias folded into the two kernels that read y1
        stats1 = ops.prelu_stats(y1, self.prelu1.weight, self.conv1x1.bias)
        z, stats2 = ops.tcn_mid(y1, self.prelu1.weight, stats1, self.norm1.weight, self.norm1.bias,
                                self.depthwise_conv.weight, self.depthwise_conv.bias, self.dilation, self.prelu2.weight,
                                self.conv1x1.bias)
        w2f, u, vb = self._folded()
        o = _conv1x1_f32(w2f, z)
        return ops.tcn_out(o, x, z.shape[1], stats2, u, vb)


class SequenceModel(_NoDerivedState, nn.Module):
    """sequence_model.py:5-123.  "TCN": runnable module.  "LSTM": parameter container whose forward goes through
    the hand-written kernels (input must already be packed time-major by ops.subband_pack)."""

    def __init__(self, input_size, output_size, hidden_size, num_layers, bidirectional, sequence_model="GRU",
                 output_activate_function="Tanh"):
        super().__init__()
        self.sequence_model_type = sequence_model
        if sequence_model == "LSTM":
            if bidirectional or num_layers != 2:
                raise NotImplementedError("only the 2-layer unidirectional LSTM of FullSubNet+ is built")
            self.sequence_model = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                                          batch_first=True, bidirectional=False)
            self.fc_output_layer = nn.Linear(hidden_size, output_size)
        elif sequence_model == "TCN":
            self.sequence_model = nn.Sequential(*[TCNBlock(input_size, 512, input_size, dilation=d) for d in TCN_DILATIONS],
                                                nn.ReLU())
            self.fc_output_layer = nn.Linear(input_size, output_size)
        else:
            raise NotImplementedError(f"Not implemented {sequence_model}")
        if output_activate_function:
            if output_activate_function == "ReLU":
                self.activate_function = nn.ReLU()
            elif output_activate_function == "Tanh":
                self.activate_function = nn.Tanh()
            elif output_activate_function == "ReLU6":
                self.activate_function = nn.ReLU6()
            else:
                raise NotImplementedError(f"Not implemented activation function {output_activate_function}")
        self.output_activate_function = output_activate_function
        self._plan = None
        self._plan_key = None

    # ---- TCN path ----
    use_tc_convs = False   # set by the owning model when lstm_impl == "tc": 1x1 convolutions on the tcgen05 GEMM (row N2)

    # split-precision 1x1 convolutions (tcn_cl.cu): hi + lo fp16 halves of the residual stream and of the conv1x1 / fc weights,
    # fp32 fc output.  NPPC_TCN_SPLIT=0 selects the plain fp16 operands (~1 ms faster per step at B = 64, 3-10x larger error on
    # cancelling inputs; profiles/r02_parity_report.md).
    tc_split = True

    @staticmethod
    def _split_w(w32, rows, Kp, dev):
        """fp32 [n, C] -> fp16 [rows, 3 Kp] = [Whi | Whi | Wlo] (zero padded)."""
        n, C = w32.shape
        w32 = w32.clamp(-65504, 65504)
        hi = w32.half()
        lo = (w32 - hi.float()).half()
        out = torch.zeros(rows, 3 * Kp, device=dev, dtype=torch.float16)
        out[:n, :C] = hi
        out[:n, Kp:Kp + C] = hi
        out[:n, 2 * Kp:2 * Kp + C] = lo
        return out

    def _tcn_plan(self):
        """fp16 / padded copies of the 1x1-conv weights for the channel-last tcgen05 path (derived cache, never serialised)."""
        blocks = [m for m in self.sequence_model if isinstance(m, TCNBlock)]
        ps = [p for blk in blocks for p in (blk.conv1x1.weight, blk.sconv.weight, blk.sconv.bias, blk.norm2.weight, blk.norm2.bias)]
        ps += [self.fc_output_layer.weight, self.fc_output_layer.bias]
        key = tuple((p.data_ptr(), p._version) for p in ps) + (self.tc_split,)
        if getattr(self, "_tplan", None) is None or key != self._tplan_key:
            C = blocks[0].conv1x1.weight.shape[1]
            Kp, Np = -(-C // 64) * 64, -(-C // 128) * 128
            dev = blocks[0].conv1x1.weight.device
            split = self.tc_split
            with torch.no_grad():
                plan = dict(C=C, Kp=Kp, Np=Np, blocks=[], split=split)
                for blk in blocks:
                    if split:
                        w1 = self._split_w(blk.conv1x1.weight[:, :, 0], 512, Kp, dev)
                    else:
                        w1 = torch.zeros(512, Kp, device=dev, dtype=torch.float16)
                        w1[:, :C] = blk.conv1x1.weight[:, :, 0].clamp(-65504, 65504).half()
                    w2f, u, vb = blk._folded()
                    w2 = torch.zeros(Np, 512, device=dev, dtype=torch.float16)
                    w2[:C] = w2f[:, :, 0].clamp(-65504, 65504).half()
                    plan["blocks"].append((w1, w2, u, vb))
                O = self.fc_output_layer.weight.shape[0]
                Op = -(-O // 128) * 128
                if split:
                    wfc = self._split_w(self.fc_output_layer.weight, Op, Kp, dev)
                else:
                    wfc = torch.zeros(Op, Kp, device=dev, dtype=torch.float16)
                    wfc[:O, :C] = self.fc_output_layer.weight.clamp(-65504, 65504).half()
                plan.update(O=O, Op=Op, wfc=wfc)
            self._tplan, self._tplan_key = plan, key
        return self._tplan

    def forward_tc(self, x):
        """x [B,C,T'] f32 -> [B,O,T'] f32: the whole stack channel-last, every 1x1 convolution and the output Linear on the
        tcgen05 fp16 GEMM (fp32 accumulate, fp32 residual stream), everything between them in three fused kernels per block."""
        pl = self._tcn_plan()
        B, C, T = x.shape
        M = B * T
        dev = x.device
        split = pl["split"]
        x32 = torch.empty(M, pl["Kp"], device=dev, dtype=torch.float32)   # residual stream, rows padded like xh
        xh = torch.empty(M, pl["Kp"] * (2 if split else 1), device=dev, dtype=torch.float16)   # K padding zeroed by the pack kernel
        # per-sample fp16 scale: the real / imag streams are normalised by a cancelling mean and can be huge (see tcn_cl.cu)
        x = x.contiguous()
        scale, inv_scale = ops.tcn_cl_scale(x)
        ops.tcn_cl_pack(x, pl["Kp"], inv_scale, x32, xh, split)
        blocks = [m for m in self.sequence_model if isinstance(m, TCNBlock)]
        for i, (blk, (w1, w2, u, vb)) in enumerate(zip(blocks, pl["blocks"])):
            y1 = ops.gemm_f16_tn_ex(xh, w1)
            stats1 = ops.prelu_stats_cl(y1, B, T, scale, blk.conv1x1.bias, blk.prelu1.weight)
            z, stats2 = ops.tcn_mid_cl(y1, B, T, scale, blk.conv1x1.bias, blk.prelu1.weight, stats1, blk.norm1.weight, blk.norm1.bias,
                                       blk.depthwise_conv.weight, blk.depthwise_conv.bias, blk.dilation, blk.prelu2.weight)
            o = ops.gemm_f16_tn(z, w2)
            ops.tcn_out_cl(o, x32, B, T, C, pl["Np"], pl["Kp"], stats2, u, vb, inv_scale, xh, relu_h=(i == len(blocks) - 1), split=split)
        o = ops.gemm_f16_tn_ex(xh, pl["wfc"], out_f32=split)
        relu = {"ReLU": 1, None: 0, "": 0, False: 0}.get(self.output_activate_function, None)
        if relu is None:
            raise NotImplementedError("tcgen05 TCN path: only ReLU / no output activation (every reference config)")
        return ops.tcn_cl_unpack(o, B, pl["O"], T, pl["Op"], scale, self.fc_output_layer.bias, relu)

    def forward(self, x):
        if self.sequence_model_type != "TCN":
            raise RuntimeError("the LSTM SequenceModel runs through lstm_forward(xs) on packed input")
        if self.use_tc_convs and x.is_cuda and not torch.is_grad_enabled():
            return self.forward_tc(x)
        o = self.fc_output_layer(self.sequence_model(x).permute(0, 2, 1))
        if self.output_activate_function:
            o = self.activate_function(o)
        return o.permute(0, 2, 1)

    # ---- LSTM path ----
    def _lstm_plan(self):
        ps = [getattr(self.sequence_model, f"{k}_l{l}") for l in (0, 1) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        ps += [self.fc_output_layer.weight, self.fc_output_layer.bias]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._plan is None or key != self._plan_key:
            l0, l1 = ps[0:4], ps[4:8]
            self._plan = ops.LstmPlan(l0[0], l0[1], l0[2], l0[3], l1[0], l1[1], l1[2], l1[3], ps[8], ps[9])
            self._plan_key = key
        return self._plan

    def lstm_params(self):
        """The ten tensors of the 2-layer LSTM + fc in the order the stepwise kernels take them."""
        m = self.sequence_model
        return [getattr(m, f"{k}_l{l}") for l in (0, 1) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")] + \
            [self.fc_output_layer.weight, self.fc_output_layer.bias]

    def lstm_forward(self, xs, impl: int, R: int = None):
        """xs [T', R_stride, KP] -> [R, O, T'] (the layout SequenceModel.forward returns, sequence_model.py:122).
        impl 0: fp32 SIMT; 1: fp16 tensor cores (persistent CTA-pair kernel for H = 384, the stepwise kernel for every other
        hidden size, e.g. FullSubNet's 257 -> 512 full-band LSTM); 2: split-precision stepwise tensor-core kernel (fp32 xs)."""
        if self.output_activate_function:
            raise NotImplementedError("sb_output_activate_function is False in every reference config")
        H = self.sequence_model.hidden_size
        if impl == 2:
            return ops.lstm_step_forward(self.lstm_params(), xs, R, precise=True)[0]
        if impl == 1 and H != 384:
            return ops.lstm_step_forward(self.lstm_params(), xs, R)[0]
        return self._lstm_plan().forward(xs, impl, R)
