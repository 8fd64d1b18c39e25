"""The training driver around the M-B hot path, with what the reference's driver lacks: an actual resume.

Mirrors ``train_improved_minicausal_vad`` (avenue_training_script2.py:339-468): same loop (train epoch, evaluation every 5th epoch and at
the end, best model by ``score_range``, periodic checkpoint, history JSON after every epoch), same on-disk formats --

  best_improved_model.pth       {'model_state_dict', 'optimizer_state_dict', 'epoch', 'eval_metrics'}                    s2:437-443
  checkpoint_epoch_<e>.pth      {'model_state_dict', 'optimizer_state_dict', 'scheduler_state_dict', 'epoch',
                                 'training_history'}  (+ 'rng_state', 'best_score_range': extra keys the reference ignores) s2:448-455
  improved_training_history.json  {'train_losses', 'loss_components', 'evaluation_metrics', 'epochs', 'learning_rates'}  s2:380-386, 458-459

-- and the JSON helpers of json_utils.py:5-63 (numpy scalars / arrays -> plain JSON).  The reference writes scheduler state into its
checkpoints but never reads any of it back; ``resume=`` here restores model, AdamW moments and step counters, the ReduceLROnPlateau state,
the history, the best score so far and the device RNG state, and continues with the next epoch.  Checkpoints are written by a background
thread from a host copy of the state (``AsyncCheckpointWriter``), so the device keeps training while the file is serialised.
"""
from __future__ import annotations

import json
import os
import queue
import threading
from pathlib import Path

import numpy as np
import torch

HISTORY_KEYS = ("train_losses", "loss_components", "evaluation_metrics", "epochs", "learning_rates")


def convert_to_json_serializable(obj):
    """json_utils.py:5-21 / s2:303-317."""
    if isinstance(obj, np.floating):
        return float(obj)
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, dict):
        return {key: convert_to_json_serializable(value) for key, value in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [convert_to_json_serializable(item) for item in obj]
    return obj


def safe_json_save(data, filepath, verbose=False) -> bool:
    """json_utils.py:23-43: never raises; the file appears atomically (written next to the target, then renamed)."""
    try:
        filepath = Path(filepath)
        filepath.parent.mkdir(parents=True, exist_ok=True)
        tmp = filepath.with_suffix(filepath.suffix + ".tmp")
        with open(tmp, "w") as f:
            json.dump(convert_to_json_serializable(data), f, indent=2)
        os.replace(tmp, filepath)
        if verbose:
            print(f"Data saved to {filepath}")
        return True
    except Exception as e:      # noqa: BLE001 -- the reference's contract: report and carry on
        print(f"Failed to save JSON to {filepath}: {e}")
        return False


def safe_json_load(filepath):
    """json_utils.py:45-63."""
    try:
        with open(filepath) as f:
            return json.load(f)
    except Exception as e:      # noqa: BLE001
        print(f"Failed to load JSON from {filepath}: {e}")
        return None


def _to_host(obj):
    if torch.is_tensor(obj):
        return obj.detach().to("cpu", copy=True)
    if isinstance(obj, dict):
        return {k: _to_host(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_host(v) for v in obj)
    return obj


class AsyncCheckpointWriter:
    """``save(state, path)`` snapshots the state to host memory on the caller's thread (one device->host copy, ordered after the step
    that produced it) and hands serialisation + disk I/O to a worker thread; ``wait()`` drains the queue (call before reading a file back
    or exiting).  Files appear atomically (tmp + rename), so a killed run never leaves a truncated checkpoint."""

    def __init__(self):
        self._q = queue.Queue()
        self._err = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                self._q.task_done()
                return
            state, path = item
            try:
                tmp = str(path) + ".tmp"
                torch.save(state, tmp)
                os.replace(tmp, str(path))
            except Exception as e:      # noqa: BLE001
                self._err = e
            self._q.task_done()

    def save(self, state, path):
        self._q.put((_to_host(state), path))

    def wait(self):
        self._q.join()
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def close(self):
        self._q.put(None)
        self._t.join()


def checkpoint_state(model, epoch, training_history, best_score_range):
    """The dict of s2:448-455 plus what a faithful resume needs."""
    dev = model.device
    return {"model_state_dict": model.model.state_dict(), "optimizer_state_dict": model.optimizer.state_dict(),
            "scheduler_state_dict": model.scheduler.state_dict(), "epoch": epoch, "training_history": training_history,
            "best_score_range": best_score_range,
            "rng_state": {"cuda": torch.cuda.get_rng_state(dev), "cpu": torch.get_rng_state()}}


def resume_from(model, path):
    """Restore everything ``checkpoint_state`` wrote; returns (next epoch, training_history, best_score_range)."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    model.model.load_state_dict(ck["model_state_dict"], strict=True)
    model.optimizer.load_state_dict(ck["optimizer_state_dict"])
    if "scheduler_state_dict" in ck:
        model.scheduler.load_state_dict(ck["scheduler_state_dict"])
    rng = ck.get("rng_state")
    if rng is not None:
        torch.cuda.set_rng_state(rng["cuda"], model.device)
        torch.set_rng_state(rng["cpu"])
    hist = ck.get("training_history") or {k: [] for k in HISTORY_KEYS}
    return int(ck["epoch"]) + 1, hist, float(ck.get("best_score_range", 0.0))


def train_improved_minicausal_vad(dataset_path, num_epochs: int = 100, batch_size: int = 4, save_interval: int = 20,
                                  output_dir="improved_avenue_results", resume=None, device="cuda", loaders=None, dp=None, verbose=True,
                                  on_epoch_end=None):
    """s2:339-468.  Returns (model, training_history).  ``resume``: a ``checkpoint_epoch_<e>.pth`` written by this function (or by the
    reference: then only model / optimizer / scheduler / history are restored).  ``loaders``: (train_loader, test_loader) to use instead of
    ``create_avenue_dataloaders(dataset_path, ...)``.  ``on_epoch_end(epoch, training_history)`` is called after an epoch's files are
    queued (raise from it to stop a run: queued checkpoints are still flushed)."""
    from .mb import ImprovedMiniCausalVAD

    def say(*a):
        if verbose:
            print(*a)

    if loaders is None:
        from avenue_dataset_usage import create_avenue_dataloaders
        train_loader, test_loader = create_avenue_dataloaders(dataset_path=dataset_path, batch_size=batch_size, num_workers=2, clip_length=8,
                                                              frame_size=(64, 64))
    else:
        train_loader, test_loader = loaders
    model = ImprovedMiniCausalVAD(device=device, verbose=verbose, dp=dp)
    training_history = {k: [] for k in HISTORY_KEYS}
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    best_score_range, first_epoch = 0.0, 0
    if resume is not None:
        first_epoch, training_history, best_score_range = resume_from(model, resume)
        say(f"Resumed from {resume}: continuing with epoch {first_epoch + 1}")
    writer = AsyncCheckpointWriter()
    try:
        for epoch in range(first_epoch, num_epochs):
            train_loss, loss_components = model.train_epoch_improved(train_loader)
            current_lr = model.optimizer.param_groups[0]["lr"]
            say(f"Epoch {epoch + 1}/{num_epochs}  Total Loss: {train_loss:.6f}  Anomaly Loss: {loss_components['anomaly_loss']:.6f}  "
                f"Avg Edges: {loss_components['edge_count']:.1f}  Learning Rate: {current_lr:.2e}")
            training_history["train_losses"].append(train_loss)
            training_history["loss_components"].append(loss_components)
            training_history["epochs"].append(epoch + 1)
            training_history["learning_rates"].append(current_lr)
            if epoch % 5 == 0 or epoch == num_epochs - 1:
                _, _, eval_metrics = model.evaluate_improved(test_loader, return_arrays=False)
                training_history["evaluation_metrics"].append(eval_metrics)
                if eval_metrics["score_range"] > best_score_range:
                    best_score_range = eval_metrics["score_range"]
                    writer.save({"model_state_dict": model.model.state_dict(), "optimizer_state_dict": model.optimizer.state_dict(),
                                 "epoch": epoch, "eval_metrics": eval_metrics}, output_dir / "best_improved_model.pth")
                    say(f"Saved best model (score range: {best_score_range:.6f})")
            if epoch % save_interval == 0:
                writer.save(checkpoint_state(model, epoch, training_history, best_score_range), output_dir / f"checkpoint_epoch_{epoch}.pth")
            safe_json_save(training_history, output_dir / "improved_training_history.json")
            if on_epoch_end is not None:
                on_epoch_end(epoch, training_history)
    finally:
        writer.wait()
        writer.close()
    return model, training_history
