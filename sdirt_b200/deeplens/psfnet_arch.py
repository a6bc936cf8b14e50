"""PSF MLP (deeplens/psfnet_arch.py:32-56, 291-304).  Dense GEMM chain: left to cuBLAS via torch.nn."""
import torch
import torch.nn as nn


def initialize_weights(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_uniform_(m.weight.data)
        nn.init.constant_(m.bias.data, 0)


class MLP(nn.Module):
    """3 -> h/4 -> h -> hidden_layers x (h -> h) -> out, ReLU everywhere (psfnet_arch.py:32-56).  Layers are
    created and initialised in the reference's order so that a seeded construction yields identical weights."""

    def __init__(self, in_features, out_features, hidden_features=64, hidden_layers=3):
        super().__init__()
        self.ks = int(out_features ** 0.5)
        net = [nn.Linear(in_features, hidden_features // 4, bias=True), nn.ReLU(inplace=True),
               nn.Linear(hidden_features // 4, hidden_features, bias=True), nn.ReLU(inplace=True)]
        for _ in range(hidden_layers):
            net += [nn.Linear(hidden_features, hidden_features, bias=True), nn.ReLU(inplace=True)]
        net += [nn.Linear(hidden_features, out_features, bias=True), nn.ReLU()]
        self.net = nn.Sequential(*net)
        self.net.apply(initialize_weights)

    def forward(self, inp):
        with torch.autocast(device_type="cuda", enabled=inp.is_cuda):      # @autocast() in the reference
            x = self.net(inp)
        return x.reshape(*x.shape[:-1], self.ks, self.ks)
