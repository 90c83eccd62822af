"""The training step of the reference's ``libs/trainer/trainer.py`` on the native engines.

``train_step`` is the loop body of ``trainer.train`` [trainer.py:183-198] -- ``optim.zero_grad()``,
``prediction = model(data)``, ``loss = loss_func(prediction, target, weights, meta)``, ``loss.backward()``,
``optim.step()`` -- with the model's train-mode forward / backward running in the native HC training engine
(``egn_hrnet_forward_train`` / ``egn_hrnet_backward``), the heat-map loss in ``egn_mse_hm_fwd_bwd`` and the
update in ``egn_adam_step`` / ``egn_sgd_step``.  Data loading, logging, plotting, evaluation and checkpoint
rotation of upstream's epoch loop are host orchestration outside the hot path and stay with the caller.
"""
import torch
import torch.distributed as dist


def allreduce_gradients(model, group=None):
    """Data-parallel training across ranks (the reference wraps the model in ``torch.nn.DataParallel``,
    train_IGRs.py:59: one process, the batch split over the GPUs, gradients summed; BatchNorm statistics stay per
    replica, no SyncBN).  Here: one process per GPU, every rank runs the step on its shard and the gradients are
    averaged with ONE all-reduce over the engine's flat gradient buffer (the ``.grad`` tensors are views of it),
    256 MB over NVLink instead of one collective per parameter.  Returns the number of collectives issued."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    inner = model.module if hasattr(model, 'module') else model
    tr = getattr(inner, '_train', None)
    flat = tr.get('last_grads') if tr else None
    if flat is not None:
        dist.all_reduce(flat, group=group)
        flat.div_(world)
        return 1
    n = 0
    for p in inner.parameters():                 # modules without a flat buffer: per-parameter fallback
        if p.grad is not None:
            dist.all_reduce(p.grad, group=group)
            p.grad.div_(world)
            n += 1
    return n


def train_step(model, loss_func, optim, data, target, weights, meta=None):
    """One optimisation step; returns the loss tensor (device scalar, no sync).  Under an initialised
    ``torch.distributed`` group the gradients are averaged over the ranks before the update."""
    optim.zero_grad()
    prediction = model(data)
    loss = loss_func(prediction, target, weights, meta)
    loss.backward()
    allreduce_gradients(model)
    optim.step()
    return loss.detach()


def train_epoch(model, loss_func, optim, sche, loader, report_every=0, report=None):
    """One epoch in upstream's order [trainer.py:164-198]: ``model.train()``, ``sche.step()``, then every batch."""
    model.train()
    sche.step()
    for batch_idx, (data, target, weights, meta) in enumerate(loader):
        data, target, weights = data.cuda(non_blocking=True), target.cuda(non_blocking=True), weights.cuda(non_blocking=True)
        loss = train_step(model, loss_func, optim, data, target, weights, meta)
        if report is not None and report_every and batch_idx % report_every == 0:
            report(batch_idx, float(loss))
