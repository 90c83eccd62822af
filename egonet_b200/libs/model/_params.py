"""Parameter containers that reproduce a reference ``state_dict`` from a flat key list.

The native engines are the single source of truth for the parameter
inventory (names, shapes, order: ``egn_hrnet_weight_key`` / ``egn_hrnet_weight_shape``).
``ParamTree`` turns that list into nested ``nn.Module`` containers whose
``state_dict()`` has exactly those keys in that order, so checkpoints written by
the reference (``HC.pth`` / ``L.pth``) load with ``strict=True``.
"""
import math

import torch
import torch.nn as nn

_BUFFER_LEAVES = ('running_mean', 'running_var', 'num_batches_tracked')


class ParamTree(nn.Module):
    """A bare container: children, parameters and buffers are attached by name."""

    def extra_repr(self):
        own = [n for n, _ in self.named_parameters(recurse=False)]
        return ', '.join(own)


def attach(root, key, shape):
    """Create ``root.<key>`` (dotted) as a Parameter or buffer with default init."""
    parts = key.split('.')
    node = root
    for name in parts[:-1]:
        if name not in node._modules:
            node.add_module(name, ParamTree())
        node = node._modules[name]
    leaf = parts[-1]
    if leaf == 'num_batches_tracked':
        node.register_buffer(leaf, torch.tensor(0, dtype=torch.long))
    elif leaf in _BUFFER_LEAVES:
        node.register_buffer(leaf, torch.zeros(shape) if leaf == 'running_mean' else torch.ones(shape))
    else:
        node.register_parameter(leaf, nn.Parameter(torch.zeros(shape)))
    return node


def is_norm(node):
    return 'running_mean' in node._buffers


def default_init(root):
    """torch.nn defaults: kaiming-uniform(a=sqrt(5)) conv/linear weights, uniform
    bias of bound 1/sqrt(fan_in), BatchNorm weight 1 / bias 0."""
    with torch.no_grad():
        for node in root.modules():
            w = node._parameters.get('weight')
            if w is None:
                continue
            b = node._parameters.get('bias')
            if is_norm(node):
                w.fill_(1.0)
                if b is not None:
                    b.zero_()
                continue
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            if b is not None:
                fan_in = w[0].numel()
                bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
                b.uniform_(-bound, bound)


def host_state(root):
    """(key, contiguous fp32 CPU tensor) for every floating-point entry of the state dict."""
    out = []
    for k, v in root.state_dict().items():
        if not v.dtype.is_floating_point:
            continue
        out.append((k, v.detach().to(device='cpu', dtype=torch.float32).contiguous()))
    return out
