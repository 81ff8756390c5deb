"""CPU oracle for the LogicRNNLSTM temporal head.  TEST INFRASTRUCTURE, NOT PRODUCT.

fp32 restatement of /root/reference/src/RNNModel.py driven by a state_dict: LogicCell.forward (:21-41),
LogicRNNLSTM.forward (:81-133) including its quirks (single (h,c) threaded through the layers, sort by length
without un-sorting, mask applied to outputs only, softmax over all T).  Pinned by tests/golden/rnn_ref_seed0.npz,
frozen from the UNMODIFIED reference class by oracle/make_golden_rnn.py."""
import torch
import torch.nn.functional as F


def logic_cell(sd, p, x, h, c):
    z = torch.cat((x, h), dim=1)
    lin = lambda n, v: F.linear(v, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"])
    and_o, or_o, not_o = torch.sigmoid(lin("and_gate", z)), torch.sigmoid(lin("or_gate", z)), torch.tanh(lin("not_gate", h))
    cell_new = torch.sigmoid(lin("forget_gate", z)) * c + torch.sigmoid(lin("input_gate", z)) * torch.tanh(lin("cell_gate", z))
    cell_logic = and_o * cell_new + or_o * not_o
    return torch.sigmoid(lin("output_gate", z)) * torch.tanh(cell_logic), cell_logic


def rnn_forward(sd, x, lengths=None, num_layers=2):
    B, T, _ = x.shape
    H = sd["logic_cells.0.not_gate.weight"].shape[0]
    if lengths is not None:
        lengths, idx = lengths.sort(0, descending=True)
        x = x[idx]
    h, c = torch.zeros(B, H), torch.zeros(B, H)
    outs = []
    for t in range(T):
        ht, ct = h, c
        for i in range(num_layers):
            ht, ct = logic_cell(sd, f"logic_cells.{i}", x[:, t, :] if i == 0 else ht, ht, ct)
        outs.append(ht)
        h, c = ht, ct
    outs = torch.stack(outs, dim=1)
    if lengths is not None:
        outs = outs * (torch.arange(T).expand(B, T) < lengths.unsqueeze(1)).float().unsqueeze(-1)
    a = F.linear(torch.tanh(F.linear(outs, sd["attention.0.weight"], sd["attention.0.bias"])), sd["attention.2.weight"], sd["attention.2.bias"])
    ctx = torch.sum(torch.softmax(a, dim=1) * outs, dim=1)
    y = F.linear(F.relu(F.linear(ctx, sd["classifier.0.weight"], sd["classifier.0.bias"])), sd["classifier.3.weight"], sd["classifier.3.bias"])
    return torch.sigmoid(y)
