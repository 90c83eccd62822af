"""The training step of the reference's ``libs/trainer/trainer.py`` on the native engines.

``train_step`` is the loop body of ``trainer.train`` [trainer.py:183-198] -- ``optim.zero_grad()``,
``prediction = model(data)``, ``loss = loss_func(prediction, target, weights, meta)``, ``loss.backward()``,
``optim.step()`` -- with the model's train-mode forward / backward running in the native HC training engine
(``egn_hrnet_forward_train`` / ``egn_hrnet_backward``), the heat-map loss in ``egn_mse_hm_fwd_bwd`` and the
update in ``egn_adam_step`` / ``egn_sgd_step``.  Data loading, logging, plotting, evaluation and checkpoint
rotation of upstream's epoch loop are host orchestration outside the hot path and stay with the caller.
"""
import torch


def train_step(model, loss_func, optim, data, target, weights, meta=None):
    """One optimisation step; returns the loss tensor (device scalar, no sync)."""
    optim.zero_grad()
    prediction = model(data)
    loss = loss_func(prediction, target, weights, meta)
    loss.backward()
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
