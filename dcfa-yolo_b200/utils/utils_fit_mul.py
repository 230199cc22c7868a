"""The validation half of the reference's epoch loop (utils/utils_fit_mul.py:60-95): eval-mode forward of every
validation batch under no_grad, the criterion on its outputs, the running mean the reference prints and logs as
`val_loss`.  Both steps are the drop-in's CUDA path (YoloBody.forward, nets.yolo_training.Loss -> dcfa_yolo_loss).

The reference reads the loss back with `.item()` after every batch (one host synchronisation per step); here the sum is
kept on the device in float64 -- the same arithmetic as Python's float accumulation of float32 values -- and read once.
The training half (:10-58: backward, clipping, optimiser, EMA) is out of scope."""
import torch


def validate_one_epoch(model_eval, yolo_loss, gen_val, epoch_step_val, cuda=True, local_rank=0, on_step=None):
    """Returns val_loss / epoch_step_val exactly as the reference accumulates it (:78-94).  `gen_val` yields
    (images_rgb, images_nir, bboxes) like the reference's DataLoader; `on_step(iteration, running_sum_tensor)` is an
    optional progress hook (the reference's tqdm postfix) that may, but need not, synchronise."""
    if not cuda:
        raise RuntimeError("validate_one_epoch: the drop-in runs on CUDA only (there is no CPU path)")
    model_eval.eval()
    dev = torch.device("cuda", local_rank)
    total = torch.zeros((), dtype=torch.float64, device=dev)
    steps = 0
    with torch.no_grad():
        for iteration, batch in enumerate(gen_val):
            if iteration >= epoch_step_val:
                break
            images_rgb, images_nir, bboxes = batch
            images_rgb = images_rgb.to(dev, non_blocking=True)
            images_nir = images_nir.to(dev, non_blocking=True)
            outputs = model_eval(images_rgb, images_nir)
            total += yolo_loss(outputs, bboxes).double()
            steps += 1
            if on_step is not None:
                on_step(iteration, total)
    return float(total) / max(epoch_step_val, 1) if steps else 0.0


def fit_one_epoch(*args, **kwargs):
    raise NotImplementedError("fit_one_epoch: the training half (backward, optimiser, EMA) is outside this drop-in's scope; "
                              "use validate_one_epoch for the validation half (reference utils/utils_fit_mul.py:60-95)")
