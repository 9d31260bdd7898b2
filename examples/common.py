"""Shared pieces of the parameter-learning examples: variance-normalised loss and the two optimisation loops."""
import numpy as np
import torch
from torch.utils.data import DataLoader


def nmse(prediction, target, variance):
    """Mean squared error with every output column scaled by the variance of its target."""
    return (((prediction - target) ** 2) / variance).mean()


def fit_minibatch(params, dataset, step_loss, n_epochs, batch_size=100, lr=1e-2, log=print):
    """Adam over mini-batches of `dataset`; `step_loss(batch)` returns the scalar loss of one batch.
    Returns the mean loss of every epoch."""
    opt = torch.optim.Adam(params, lr=lr)
    loader = DataLoader(dataset=dataset, batch_size=batch_size, shuffle=False)
    per_epoch = []
    for epoch in range(n_epochs):
        seen = []
        for batch in loader:
            opt.zero_grad()
            value = step_loss(batch)
            value.backward()
            opt.step()
            seen.append(float(value.detach()))
        per_epoch.append(float(np.mean(seen)))
        log(f"i: {epoch} loss: {per_epoch[-1]}")
    return per_epoch


def fit_full_batch(params, loss_of_iteration, n_iters, lr=1e-3, every=100, before_step=None, log=print):
    """Adam on one full batch; `loss_of_iteration()` returns the scalar loss, `before_step(i)` may freeze / unfreeze
    parameters.  Returns the loss of every iteration."""
    opt = torch.optim.Adam(params, lr=lr)
    trace = []
    for i in range(n_iters):
        opt.zero_grad()
        value = loss_of_iteration()
        trace.append(float(value.detach()))
        if i % every == 0:
            log(f"i: {i}, loss: {trace[-1]}")
        if before_step is not None:
            before_step(i)
        value.backward()
        opt.step()
    return trace
