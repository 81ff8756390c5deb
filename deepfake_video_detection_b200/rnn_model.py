"""Drop-in for the reference's `src/RNNModel.py` (`LogicCell`, `LogicRNNLSTM`, `create_model`).

Same constructor, parameter names (state_dict schema), `forward(x, lengths=None) -> (B,1)` sigmoid probabilities
and `predict`.  In eval mode the recurrence runs in libdfd_b200.so (one tcgen05 GEMM + one fused cell kernel per
step and layer, csrc/rnn.cu); training mode keeps an eager graph for autograd.  No CPU path for inference."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .engine import DEFAULT_PRECISION, PRECISIONS, _stream_ptr


class LogicCell(nn.Module):                     # RNNModel.py:5-41 (parameter containers + eager training forward)
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.and_gate = nn.Linear(input_size + hidden_size, hidden_size)
        self.or_gate = nn.Linear(input_size + hidden_size, hidden_size)
        self.not_gate = nn.Linear(hidden_size, hidden_size)
        self.forget_gate = nn.Linear(input_size + hidden_size, hidden_size)
        self.input_gate = nn.Linear(input_size + hidden_size, hidden_size)
        self.cell_gate = nn.Linear(input_size + hidden_size, hidden_size)
        self.output_gate = nn.Linear(input_size + hidden_size, hidden_size)

    def forward(self, x, hidden, cell):
        z = torch.cat((x, hidden), dim=1)
        cell_new = torch.sigmoid(self.forget_gate(z)) * cell + torch.sigmoid(self.input_gate(z)) * torch.tanh(self.cell_gate(z))
        cell_logic = torch.sigmoid(self.and_gate(z)) * cell_new + torch.sigmoid(self.or_gate(z)) * torch.tanh(self.not_gate(hidden))
        return torch.sigmoid(self.output_gate(z)) * torch.tanh(cell_logic), cell_logic


class LogicRNNLSTM(nn.Module):                  # RNNModel.py:43-147
    def __init__(self, input_size=1024, hidden_size=512, num_layers=2, dropout=0.5, precision: str = DEFAULT_PRECISION):
        super().__init__()
        self.hidden_size, self.num_layers, self.input_size, self.precision = hidden_size, num_layers, input_size, precision
        self.logic_cells = nn.ModuleList([LogicCell(input_size if i == 0 else hidden_size, hidden_size) for i in range(num_layers)])
        self.dropout = nn.Dropout(dropout)
        self.attention = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.Tanh(), nn.Linear(hidden_size, 1), nn.Softmax(dim=1))
        self.classifier = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(), nn.Dropout(dropout), nn.Linear(hidden_size, 1))
        self._handle, self._key = None, None

    def _pack(self, device):
        tensors = list(self.state_dict(keep_vars=True).items())
        key = (str(device), self.precision, tuple((t.data_ptr(), t._version) for _, t in tensors))
        if self._handle is not None and key == self._key:
            return self._handle
        lib = _lib.load()
        if self._handle is not None:
            lib.dfd_rnn_free_weights(self._handle)
        keep = [(k.encode(), v.detach().to("cpu", torch.float32).contiguous()) for k, v in tensors]
        n = len(keep)
        names = (C.c_char_p * n)(*[k for k, _ in keep])
        data = (C.c_void_p * n)(*[v.data_ptr() for _, v in keep])
        numel = (C.c_int64 * n)(*[v.numel() for _, v in keep])
        h = C.c_void_p()
        with torch.cuda.device(device):
            rc = lib.dfd_rnn_pack_weights(n, names, data, numel, self.input_size, self.hidden_size, self.num_layers,
                                          PRECISIONS[self.precision], C.byref(h))
        if rc:
            raise RuntimeError(f"dfd_rnn_pack_weights failed ({rc}): {lib.dfd_rnn_last_error().decode()}")
        self._handle, self._key = h, key
        return h

    def __del__(self):                    # release the packed device weights with the module
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().dfd_rnn_free_weights(self._handle)
                self._handle = None
        except Exception:
            pass

    def forward(self, x, lengths=None):
        batch_size, seq_length, _ = x.size()
        if lengths is not None:                                   # :92-95 (outputs stay in sorted order, as in the reference)
            lengths, sort_idx = lengths.sort(0, descending=True)
            x = x[sort_idx]
        if self.training:
            return self._forward_eager(x, lengths)
        if x.device.type != "cuda":
            raise RuntimeError("LogicRNNLSTM (B200 build): inference needs a CUDA tensor; there is no CPU fallback")
        lib = _lib.load()
        h = self._pack(x.device)
        x = x.contiguous().float()
        ln = lengths.to(device=x.device, dtype=torch.int32).contiguous() if lengths is not None else None
        prob = torch.empty((batch_size, 1), dtype=torch.float32, device=x.device)
        nbytes = C.c_size_t()
        if lib.dfd_rnn_workspace_bytes(h, batch_size, seq_length, C.byref(nbytes)):
            raise RuntimeError(lib.dfd_rnn_last_error().decode())
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.dfd_rnn_forward(h, x.data_ptr(), None if ln is None else ln.data_ptr(), batch_size, seq_length,
                                     prob.data_ptr(), ws.data_ptr(), nbytes.value, _stream_ptr(x.device))
        if rc:
            raise RuntimeError(f"dfd_rnn_forward failed ({rc}): {lib.dfd_rnn_last_error().decode()}")
        return prob

    def _forward_eager(self, x, lengths):
        b, t, _ = x.size()
        h = torch.zeros(b, self.hidden_size, device=x.device)
        c = torch.zeros(b, self.hidden_size, device=x.device)
        outs = []
        for step in range(t):
            ht, ct = h, c
            for i, cell in enumerate(self.logic_cells):
                ht, ct = cell(x[:, step, :] if i == 0 else ht, ht, ct)
                if i < self.num_layers - 1:
                    ht = self.dropout(ht)
            outs.append(ht)
            h, c = ht, ct
        outs = torch.stack(outs, dim=1)
        if lengths is not None:
            mask = (torch.arange(t, device=x.device).expand(b, t) < lengths.unsqueeze(1)).float().unsqueeze(-1)
            outs = outs * mask
        ctx = torch.sum(self.attention(outs) * outs, dim=1)
        return torch.sigmoid(self.classifier(ctx))

    def predict(self, x, lengths=None):
        with torch.no_grad():
            return (self.forward(x, lengths) >= 0.5).float()


def create_model(config=None):                  # RNNModel.py:149-170
    config = config or {"input_size": 1024, "hidden_size": 512, "num_layers": 2, "dropout": 0.5}
    return LogicRNNLSTM(input_size=config.get("input_size", 1024), hidden_size=config.get("hidden_size", 512),
                        num_layers=config.get("num_layers", 2), dropout=config.get("dropout", 0.5))
