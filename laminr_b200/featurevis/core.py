"""Generic gradient ascent for an arbitrary differentiable `f` (reference featurevis/core.py:9-136).

This is the compatibility path for user-supplied response models (stock autograd).  Simulated populations never take
it: `get_mei_dict` dispatches them to the batched shared-memory-resident kernel lmr_mei_ascent."""
import warnings

import torch
from torch import optim


def gradient_ascent(f, x, transform=None, regularization=None, gradient_f=None, post_update=None, optim_name="SGD", step_size=0.1,
                    optim_kwargs={}, num_iterations=1000, save_iters=None, print_iters=100):
    if x.dtype != torch.float32:
        raise ValueError("x must be of torch.float32 dtype")
    x = x.detach().clone().requires_grad_()
    if optim_name == "SGD":
        optimizer = optim.SGD([x], lr=step_size, **optim_kwargs)
    elif optim_name == "Adam":
        optimizer = optim.Adam([x], lr=step_size, **optim_kwargs)
    else:
        raise ValueError("Expected optim_name to be 'SGD' or 'Adam'")
    fevals_dev, reg_terms, saved_xs = [], [], []
    i = 0
    for i in range(1, num_iterations + 1):
        if x.grad is not None:
            x.grad.zero_()
        tx = x if transform is None else transform(x, iteration=i)
        feval = f(tx)
        fevals_dev.append(feval.detach().reshape(()))  # kept on the device: no host sync per iteration
        reg_term = 0
        if regularization is not None:
            reg_term = regularization(tx, iteration=i)
            reg_terms.append(reg_term.item())
        (-feval + reg_term).backward()
        if x.grad is None:
            raise RuntimeError("Gradient did not reach x.")
        if gradient_f is not None:
            x.grad = gradient_f(x.grad, iteration=i)
        optimizer.step()
        if post_update is not None:
            with torch.no_grad():
                x[:] = post_update(x, iteration=i)
        if i % print_iters == 0:
            print("Iter {}: f(x) = {:.2f}, std(x) = {:.2f}".format(i, feval.item(), x.std().item()))
        if save_iters is not None and i % save_iters == 0:
            saved_xs.append(x.detach().clone())
    with torch.no_grad():
        tx = x if transform is None else transform(x, iteration=i + 1)
        fevals_dev.append(f(tx).detach().reshape(()))
        if regularization is not None:
            reg_terms.append(regularization(tx, iteration=i + 1).item())
    fevals = torch.stack(fevals_dev).cpu().tolist()
    if len(fevals) > 1 and all(v == fevals[0] for v in fevals):
        warnings.warn("f(x) never changed: gradient for x may be all zero")
    print("Final f(x) = {:.2f}".format(fevals[-1]))
    return (x.detach().clone() if save_iters is None else saved_xs), fevals, reg_terms
