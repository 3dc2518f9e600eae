"""Training / evaluation loop and ranking metrics with the reference's API (src/train.py).

`compute_HR`, `compute_NDCG`, `evaluate` and `train` keep the reference signatures and return
values; the arithmetic runs in libcarca_b200.so.  `evaluate` keeps its accumulators on the device
and reads them back once per call instead of three `.item()` syncs per batch (src/train.py:47-50).
"""
from __future__ import annotations

import os
from datetime import datetime
from typing import Optional, Tuple, Union

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler
from torch.utils.data import DataLoader

from . import ops
from .abstract import Model
from .carca import BinaryCrossEntropy
from .utils import get_mask, to


def _metric_acc(device) -> torch.Tensor:
    return torch.zeros(3, dtype=torch.float64, device=device)


def compute_HR(y_pred: torch.Tensor, y_true: torch.Tensor, k: int) -> float:
    """Number of labelled candidates ranked in the top k (src/train.py:15-21)."""
    acc = _metric_acc(y_pred.device)
    ops.rank_metrics_(acc, y_pred, y_true, k)
    return float(acc[0].item())


def compute_NDCG(y_pred: torch.Tensor, y_true: torch.Tensor, k: int) -> float:
    """Sum over top-k labelled candidates of 1/log2(rank+2) (src/train.py:24-32)."""
    acc = _metric_acc(y_pred.device)
    ops.rank_metrics_(acc, y_pred, y_true, k)
    return float(acc[1].item())


def evaluate(model: Model, loader: DataLoader, device: str, k: int,
             reduce_fn=None) -> Tuple[float, float, float]:
    """(HR@k, NDCG@k, mean batch loss) over `loader` (src/train.py:35-53).

    `reduce_fn(tensor)` (data-parallel runs) all-reduces the 5 device accumulators
    [hits, ndcg, users, loss_sum, n_batches] before the single device->host read.
    """
    model = model.eval().to(device)
    acc = torch.zeros(4, dtype=torch.float64, device=device)     # hits, ndcg, users, sum of batch losses
    n_batches = 0
    with torch.no_grad():
        for batch in loader:
            p_x, p_a, p_c, o_x, o_a, o_c, y_true = to(*batch, device=device)
            y_pred = model.forward(profile=(p_x, p_a, p_c), targets=[(o_x, o_a, o_c)])
            # BinaryCrossEntropy over get_mask(o_x) + compute_HR + compute_NDCG of the batch in one launch
            ops.eval_metrics_(acc, y_pred, y_true, o_x, k)
            n_batches += 1
    stats = torch.cat([acc, torch.tensor([float(n_batches)], dtype=torch.float64, device=device)])
    if reduce_fn is not None:
        reduce_fn(stats)
    hits, ndcg, total, lsum, nb = stats.tolist()          # the one device->host sync
    total = max(total, 1.0)
    return hits / total, ndcg / total, lsum / max(nb, 1.0)


def train(
    model: Model,
    train_loader: DataLoader,
    val_loader: DataLoader,
    test_loader: DataLoader,
    device: str,
    optim: Optimizer,
    epochs: int,
    top_k: int = 10,
    verbose: int = 1,
    early_stop: int = 10,
    datadir: str = "model",
    scheduler: Union[_LRScheduler, None] = None,
    loss_fn: Optional[BinaryCrossEntropy] = None,
    is_main: bool = True,
) -> Model:
    """Epoch loop, CSV log, save-best checkpoint and early stop as src/train.py:56-152.

    Differences, all outside the hot path: the per-step loss is accumulated on the device (one
    read per epoch instead of `loss.item()` per step, :97); the best checkpoint is re-read with
    `weights_only=False` (torch >= 2.6 rejects the reference's bare `torch.load`, :142);
    `loss_fn` / `is_main` let the data-parallel wrapper inject its loss and keep file output on
    rank 0.
    """
    if is_main:
        os.makedirs(datadir, exist_ok=True)
    loss_fn = loss_fn or BinaryCrossEntropy()
    model = model.train().to(device)
    best, no_improve = 0, 0
    start = datetime.now()
    logpath = f"{start.year}-{start.month}-{start.day}T{start.hour}-{start.minute}-{start.second}.csv"
    logfile = open(f"./{datadir}/{logpath}", "a") if is_main else None
    epoch = 0
    for epoch in range(1, epochs + 1):
        sum_loss = torch.zeros((), dtype=torch.float32, device=device)
        for i, batch in enumerate(train_loader, start=1):
            p_x, p_a, p_c, o_x, o_a, o_c, y_true = to(*batch, device=device)
            half = o_x.shape[1] // 2
            pos = (o_x[:, :half], None if o_a is None else o_a[:, :half], o_c[:, :half])
            neg = (o_x[:, half:], None if o_a is None else o_a[:, half:], o_c[:, half:])
            optim.zero_grad()
            y_pred = model.forward(profile=(p_x, p_a, p_c), targets=[pos, neg])
            loss = loss_fn.forward(y_pred, y_true, get_mask(o_x))
            loss.backward()
            optim.step()
            sum_loss += loss.detach()
            if verbose == 2 and is_main:
                now = datetime.now().strftime("%H:%M:%S")
                print(f"{now} - Batch {i:03d}: Loss = {(sum_loss.item() / i):.4f}")
        train_loss = sum_loss.item() / max(len(train_loader), 1)
        if verbose in [1, 2] and is_main:
            now = datetime.now().strftime("%H:%M:%S")
            print(f"{now} - Epoch {(epoch):03d}: Train Loss = {train_loss:.4f}")
            logfile.write(f"{now};{epoch};train;{train_loss};;\n")
        if scheduler is not None:
            scheduler.step()
        HR, NDCG, loss = evaluate(model, val_loader, device, top_k)
        model = model.train().to(device)
        if NDCG > best:
            best, no_improve = NDCG, 0
            if is_main:
                for f in [f for f in os.listdir(datadir) if f.endswith(".pth")]:
                    os.remove(os.path.join(datadir, f))
                torch.save(model, os.path.join(datadir, f"{epoch:03d}_{HR:.4f}_{NDCG:.4f}.pth"))
        else:
            no_improve += 1
        if verbose in [1, 2] and is_main:
            now = datetime.now().strftime("%H:%M:%S")
            print(f"{now} - Epoch {epoch:03d}: Val Loss = {loss:.4f} HR = {HR:.4f}, NDCG = {NDCG:.4f}")
            logfile.write(f"{now};{epoch};val;{loss};{HR};{NDCG}\n")
        if no_improve >= early_stop:
            if is_main:
                print(f"No improvement in {no_improve} epochs, early stopping...")
            break
        if logfile is not None:
            logfile.flush()
    if is_main:
        saved = [os.path.join(datadir, f) for f in os.listdir(datadir) if f.endswith(".pth")]
        if saved:
            model = torch.load(saved[0], weights_only=False)
        if test_loader is not None:
            HR, NDCG, loss = evaluate(model, test_loader, device, top_k)
            now = datetime.now().strftime("%H:%M:%S")
            print(f"{now} - Epoch {epoch:03d}: Test Loss = {loss:.4f} HR = {HR:.4f}, NDCG = {NDCG:.4f}")
            logfile.write(f"{now};{epoch};test;{loss};{HR};{NDCG}\n")
        logfile.close()
    return model
