"""B200-native drop-in for the reference's ``libs/optimizer/optimizer.py``.

``prepare_optim(model, cfgs)`` keeps the upstream signature and configuration keys
[optimizer.py:9-41] (``optim_type`` 'adam' | 'sgd', ``lr``, ``weight_decay``, ``momentum``,
``milestones``, ``gamma``) and returns ``(optimizer, scheduler)`` with ``torch.optim`` semantics.
When the model keeps its parameters in one flat device buffer (the native HC training engine does),
the update is ONE fused kernel over that buffer (``egn_adam_step`` / ``egn_sgd_step``) instead of a
launch train per parameter; otherwise the torch optimisers are returned, as upstream.
"""
import torch

from ... import _native as N


class FlatOptimizer(torch.optim.Optimizer):
    """Adam / SGD over the flat parameter buffer of a natively trained module.  ``param_groups`` (one group:
    ``lr``, ``weight_decay``, ...) behaves as in torch, so ``MultiStepLR`` drives it unchanged."""

    def __init__(self, module, kind, lr, weight_decay=0.0, momentum=0.0, betas=(0.9, 0.999), eps=1e-8):
        params = [p for p in module.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, weight_decay=weight_decay, momentum=momentum, betas=betas, eps=eps))
        self.module, self.kind, self.steps = module, kind, 0
        self.bufs = None

    def _flat(self):
        dev = next(self.module.parameters()).device
        tr = self.module._train_state(dev)
        flat = tr['flat']
        if self.bufs is None or self.bufs[0].shape != flat.shape or self.bufs[0].device != flat.device:
            mask = torch.zeros(flat.numel(), dtype=torch.uint8, device=flat.device)
            named = dict(self.module.named_parameters())
            for k, off, n, _ in tr['entries']:
                if k in named and named[k].requires_grad:
                    mask[off:off + n] = 1
            self.bufs = (torch.zeros_like(flat), torch.zeros_like(flat), mask, torch.zeros_like(flat))
        return tr, flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        tr, flat = self._flat()
        m, v, mask, grads = self.bufs
        named = dict(self.module.named_parameters())
        last = tr.get('last_grads')
        if last is not None and all(p.grad is not None and p.grad.data_ptr() == last.data_ptr() + 4 * off
                                    for k, off, n, _ in tr['entries'] for p in [named.get(k)] if p is not None and p.requires_grad):
            grads = last                                     # the .grad tensors ARE the engine's flat gradient buffer
        else:
            grads.zero_()
            for k, off, n, _ in tr['entries']:               # gather the autograd gradients into the flat layout
                p = named.get(k)
                if p is not None and p.grad is not None:
                    grads[off:off + n].copy_(p.grad.reshape(-1))
        g = self.param_groups[0]
        self.steps += 1
        with torch.cuda.device(flat.device):
            if self.kind == 'adam':
                N.check(N.lib().egn_adam_step(N.ptr(flat), N.ptr(grads), N.ptr(m), N.ptr(v), N.ptr(mask), flat.numel(),
                                              self.steps, g['lr'], g['betas'][0], g['betas'][1], g['eps'],
                                              g['weight_decay'], N.current_stream()))
            else:
                N.check(N.lib().egn_sgd_step(N.ptr(flat), N.ptr(grads), N.ptr(m), N.ptr(mask), flat.numel(), self.steps,
                                             g['lr'], g['momentum'], g['weight_decay'], 0, N.current_stream()))
        return loss


def prepare_optim(model, cfgs):
    """[optimizer.py:9-41]"""
    oc = cfgs['optimizer']
    inner = model.module if hasattr(model, 'module') else model
    native = hasattr(inner, '_train_state') and next(inner.parameters()).is_cuda
    if oc['optim_type'] not in ('adam', 'sgd'):
        raise NotImplementedError
    if native:
        optimizer = FlatOptimizer(inner, oc['optim_type'], oc['lr'], oc['weight_decay'],
                                  oc['momentum'] if oc['optim_type'] == 'sgd' else 0.0)
    elif oc['optim_type'] == 'adam':
        optimizer = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=oc['lr'],
                                     weight_decay=oc['weight_decay'])
    else:
        optimizer = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=oc['lr'],
                                    momentum=oc['momentum'], weight_decay=oc['weight_decay'])
    scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=oc['milestones'], gamma=oc['gamma'])
    return optimizer, scheduler
