"""Checkpoint files in the reference's format:

    {'epoch': int, 'state_dict': model.state_dict(), 'best_acc': float, 'optimizer': optimizer.state_dict()}

written by `save_checkpoint` (main.py:540-555: `<model_type>_checkpoint.pth.tar`, `val_` prefix for validation
checkpoints, a `<model_type>_<EE>ep_<acc>acc_checkpoint.pth.tar` copy for the best one, rank 0 only) and read at
main.py:251-265 (resume) and test.py:152-153 (inference; only 'state_dict' is required there). A file written by the
reference loads here and the other way round: the model keeps the 254-key state_dict and `TrainStep.state_dict()` has
the layout of `torch.optim.SGD.state_dict()`.
"""
import os
import shutil

import torch


def checkpoint_name(model_type, is_val=False, suffix_name='checkpoint.pth.tar'):
    if is_val:
        suffix_name = 'val_' + suffix_name
    return model_type + '_' + suffix_name, suffix_name


def save_checkpoint(state, args, is_best, is_val=False, suffix_name='checkpoint.pth.tar', directory=''):
    """Same arguments and file names as the reference (main.py:540-555). `args` needs `model_type`,
    `multiprocessing_distributed` and `rank`; ranks other than 0 write nothing. Returns the paths written."""
    distributed = bool(getattr(args, 'multiprocessing_distributed', False))
    if distributed and getattr(args, 'rank', 0) != 0:
        return []
    name, suffix_name = checkpoint_name(args.model_type, is_val, suffix_name)
    current = os.path.join(directory, name)
    torch.save(state, current)
    written = [current]
    if is_best:
        best = os.path.join(directory, args.model_type + '_{:02d}ep_'.format(state['epoch'])
                            + '{:.4f}acc_'.format(state['best_acc']) + suffix_name)
        shutil.copyfile(current, best)
        written.append(best)
    return written


def make_state(epoch, model, best_acc, optimizer):
    """The dict main.py:343-349 saves; a DDP-style wrapper (`.module`) is unwrapped the way the reference does."""
    inner = model.module if hasattr(model, 'module') else model
    return {'epoch': epoch + 1, 'state_dict': inner.state_dict(), 'best_acc': best_acc,
            'optimizer': optimizer.state_dict()}


def _strip_module_prefix(sd):
    # checkpoints saved from a wrapped model without `.module` (not the reference's own path, but common in the wild)
    if sd and all(k.startswith('module.') for k in sd):
        return {k[len('module.'):]: v for k, v in sd.items()}
    return sd


def load_checkpoint(path, model, optimizer=None, map_location='cpu', strict=True):
    """Resume (main.py:251-265) / inference load (test.py:152-153). Returns (start_epoch, best_acc); both are None-safe
    for inference-only files that hold just 'state_dict'. Raises FileNotFoundError like the reference's resume path."""
    if not os.path.isfile(path):
        raise FileNotFoundError('Valid checkpoint for resume is not found.')
    checkpoint = torch.load(path, map_location=map_location, weights_only=False)
    inner = model.module if hasattr(model, 'module') else model
    inner.load_state_dict(_strip_module_prefix(checkpoint['state_dict']), strict=strict)
    if optimizer is not None and 'optimizer' in checkpoint:
        optimizer.load_state_dict(checkpoint['optimizer'])
    return checkpoint.get('epoch', 0), checkpoint.get('best_acc', 0.0)
