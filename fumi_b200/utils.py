"""Flags, factories, AM3 metrics and checkpoints: mirror of the reference's fumi/utils/utils.py.

The flag table is the reference's (utils.py:19-229) flag for flag -- names, types, defaults -- plus
three additions that do not rename anything: --tasks_per_batch (alias of --batch_size for the
batched engine), --precision and --synthetic.  wandb is optional and never on the compute path.
"""
import argparse
import os
import shutil
import time

import numpy as np
import torch

from . import am3, fumi, maml
from .optim import FusedAdam

try:
    import wandb                      # reference: utils.py:6; logging only
except Exception:                     # pragma: no cover
    wandb = None


def parser():
    p = argparse.ArgumentParser(description="Multimodal image classification")
    # data config
    p.add_argument("--wandb_entity", type=str, default="multimodal-image-cls", help="W&B entity")
    p.add_argument("--wandb_project", type=str, default="fumi", help="W&B project")
    p.add_argument("--dataset", type=str, default="inat-anim", help="Dataset to use (inat-anim, supervised-inat-anim")
    p.add_argument("--data_dir", type=str, default="./data", help="Directory to use for data")
    p.add_argument("--checkpoint", type=str, default=None, help="Path to pretrained model")
    p.add_argument("--log_dir", type=str, default="./results", help="Directory to use for results")
    p.add_argument("--remove_stop_words", action="store_true", help="Whether to remove stop words")
    p.add_argument("--colab", action="store_true", help="Whether the script is running on Google Colab")
    # optimizer config
    p.add_argument("--epochs", type=int, default=50000, help="Number of meta-learning batches to train for")
    p.add_argument("--optim", type=str, default="adam", help="Optimiser")
    p.add_argument("--lr", type=float, default=3e-5, help="Learning rate")
    p.add_argument("--momentum", type=float, default=0.9, help="Momentum for SGD")
    p.add_argument("--batch_size", type=int, default=4, help="Number of tasks in mini-batch")
    p.add_argument("--weight_decay", type=float, default=5e-4, help="L2 regulariser")
    p.add_argument("--num_warmup_steps", type=float, default=10, help="Warm up lr scheduler")
    # dataloader config
    p.add_argument("--num_shots", type=int, default=5, help="Number of examples per class (k-shot)")
    p.add_argument("--num_ways", type=int, default=5, help="Number of classes per task (N-way)")
    p.add_argument("--num_shots_test", type=int, default=32, help="Number of examples per class in query set")
    p.add_argument("--augment", action="store_true", help="Augment data with image transformations")
    p.add_argument("--num_workers", type=int, default=0, help="Number of workers for dataloader")
    p.add_argument("--image_embedding_model", type=str, default="resnet-152",
                   help="resnet-152 embedding (2048 dimensions) or resnet-34 (512 dimensions)")
    # model config
    p.add_argument("--model", type=str, default="fumi", help="Model to be trained")
    p.add_argument("--prototype_dim", type=int, default=64, help="Dimension of latent space")
    p.add_argument("--im_encoder", type=str, default="precomputed",
                   help="Type of vision feature extractor (resnet, precomputed)")
    p.add_argument("--im_emb_dim", type=int, default=2048, help="Dimension of image embedding (if precomputed)")
    p.add_argument("--im_hid_dim", type=int, nargs="+", default=[256, 64], help="Hidden dimension of image model")
    p.add_argument("--text_encoder", type=str, choices=["glove", "w2v", "RNN", "RNNhid", "BERT", "rand"],
                   default="BERT", help="Type of text embedding (glove, w2v, RNN, RNNhid, BERT, rand)")
    p.add_argument("--pooling_strat", type=str, default="mean",
                   help="Pooling strategy if using word embeddings (mean, max)")
    p.add_argument("--fine_tune", action="store_true", help="Whether to fine tune text encoder")
    p.add_argument("--text_type", type=str, nargs="+", default=["description"],
                   help="What to use for text embedding (label, description or common_name)")
    p.add_argument("--text_emb_dim", type=int, default=768, help="Dimension of text embedding (if precomputed)")
    p.add_argument("--text_hid_dim", type=int, default=256,
                   help="Hidden dimension for NN mapping to prototypes and lamda")
    p.add_argument("--dropout", type=float, default=0.25, help="Dropout rate")
    p.add_argument("--step_size", type=float, default=0.01, help="MAML step size")
    p.add_argument("--first_order", action="store_true", help="Whether to use first-order MAML")
    p.add_argument("--num_train_adapt_steps", type=int, default=5,
                   help="Number of MAML inner train loop adaptation steps")
    p.add_argument("--num_test_adapt_steps", type=int, default=100,
                   help="Number of MAML inner test loop adaptation steps")
    p.add_argument("--init_all_layers", action="store_true",
                   help="Whether to initialise all (vs. last) layer weights in FUMI")
    p.add_argument("--norm_hypernet", action="store_true",
                   help="Whether to normalize output of the FUMI hypernetwork (tanh)")
    p.add_argument("--hypernet_bias_init", action="store_true", help="Whether to initialise hypernet bias for policy")
    p.add_argument("--lamda_fixed", default=None, type=int,
                   help="Lambda fixed for am3. Lambda = 0 is text only, Lambda = 1 is image only")
    # clip config
    p.add_argument("--clip_latent_dim", type=int, default=512, help="Dimension of CLIP latent space")
    # run config
    p.add_argument("--seed", type=int, default=123, help="patience for early stopping")
    p.add_argument("--patience", type=int, default=10000, help="Early stopping patience")
    p.add_argument("--eval_freq", type=int, default=2500, help="Number of batches between validation/checkpointing")
    p.add_argument("--wandb_experiment", type=str, default="debug", help="Name for experiment (for wandb group)")
    p.add_argument("--evaluate", action="store_true", help="skip training")
    p.add_argument("--num_ep_test", type=int, default=1000,
                   help="Number of few-shot episodes to compute test accuracy")
    p.add_argument("--disable_cuda", action="store_true", help="don't use GPU")
    p.add_argument("--wandb_offline", action="store_true", help="don't save to wandb")
    # additions of this implementation (no reference flag is renamed)
    p.add_argument("--tasks_per_batch", type=int, default=None,
                   help="tasks per meta-batch for the batched engine (overrides --batch_size)")
    p.add_argument("--precision", type=int, default=2, choices=[0, 1, 2],
                   help="dense layers: 2 (default) tcgen05, bank-sized contractions + Gram on fp16 hi/lo planes, the rest "
                        "3xTF32; 1 tcgen05 3xTF32 everywhere; 0 fp32 FMA")
    p.add_argument("--synthetic", action="store_true",
                   help="use the synthetic iNat-Anim-shaped banks instead of --data_dir")
    return p


def build_model(args, dictionary=None):
    """The nn.Module of utils.py:232-266, constructed on the CPU exactly as the reference constructs it (same
    layer creation order, so the same torch.manual_seed gives bit-identical initial weights)."""
    if args.model == "maml":
        model = maml.PureImageNetwork(im_embed_dim=args.im_emb_dim, n_way=args.num_ways, hidden_dims=args.im_hid_dim)
    elif args.model == "fumi":
        model = fumi.FUMI(n_way=args.num_ways, im_emb_dim=args.im_emb_dim, im_hid_dim=args.im_hid_dim,
                          text_encoder=args.text_encoder, text_emb_dim=args.text_emb_dim,
                          text_hid_dim=args.text_hid_dim, dropout_rate=args.dropout, dictionary=dictionary,
                          pooling_strat=args.pooling_strat, init_all_layers=args.init_all_layers,
                          norm_hypernet=args.norm_hypernet, fine_tune=args.fine_tune,
                          init_bias=args.hypernet_bias_init)
    elif args.model == "clip":
        raise NotImplementedError("--model clip is not episodic and is outside this implementation's path")
    else:
        model = am3.AM3(im_encoder=args.im_encoder, im_emb_dim=args.im_emb_dim, text_encoder=args.text_encoder,
                        text_emb_dim=args.text_emb_dim, text_hid_dim=args.text_hid_dim,
                        prototype_dim=args.prototype_dim, dropout=args.dropout, fine_tune=args.fine_tune,
                        dictionary=dictionary, pooling_strat=args.pooling_strat, lamda_fixed=args.lamda_fixed)
    return model


def init_model(args, dictionary, watch=True):
    """utils.py:232-274 (CLIP is outside the episodic path: SURVEY.md section 2 row 9)."""
    model = build_model(args, dictionary)
    model.to(args.device)
    model._get_engine(args.device).precision = int(getattr(args, "precision", 2))
    if hasattr(model, "dropout_base_seed") or args.model in ("fumi", "am3"):
        model.dropout_base_seed = int(getattr(args, "seed", 0))
    return model


def init_optim(args, model):
    """utils.py:277-299.  adam / adamw run on the fused flat-buffer kernel."""
    params = [p for p in model.parameters()]
    if args.optim == "adam":
        return FusedAdam(params, lr=args.lr, weight_decay=args.weight_decay)
    if args.optim == "SGD":
        return torch.optim.SGD(params=params, lr=args.lr, weight_decay=args.weight_decay, momentum=args.momentum)
    # the reference builds transformers.AdamW(params, lr) (utils.py:11,286-294), whose defaults are
    # weight_decay=0.0, eps=1e-6, betas=(0.9, 0.999), correct_bias=True -- not torch.optim.AdamW's 1e-2 / 1e-8
    if args.optim == "adamw":
        return FusedAdam(params, lr=args.lr, weight_decay=0.0, eps=1e-6, decoupled=True)
    if args.optim == "adamw_lin_schedule":
        opt = FusedAdam(params, lr=args.lr, weight_decay=0.0, eps=1e-6, decoupled=True)
        warm, total = args.num_warmup_steps, args.epochs

        def lr_lambda(step):                     # transformers.get_linear_schedule_with_warmup
            if step < warm:
                return float(step) / float(max(1, warm))
            return max(0.0, float(total - step) / float(max(1, total - warm)))
        return opt, torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda)
    raise NotImplementedError()


# ---- logging / checkpoints -----------------------------------------------------------------------
def run_dir(args=None):
    """wandb.run.dir when a run is active (utils.py:412); else a directory unique to this run,
    <log_dir>/run-<pid>-<start time>, remembered on `args` (the reference's wandb.run.dir is unique per run, so a
    run never sees another run's best.pth.tar)."""
    if wandb is not None and getattr(wandb, "run", None) is not None:
        return wandb.run.dir
    d = getattr(args, "_run_dir", None) if args is not None else None
    if d is None:
        base = getattr(args, "log_dir", "./results") if args is not None else "./results"
        d = os.path.join(base, f"run-{os.getpid()}-{int(time.time() * 1000)}")
        if args is not None:
            args._run_dir = d
    os.makedirs(d, exist_ok=True)
    return d


def is_main_process():
    """Rank 0 of a torchrun launch (or the only process): the one that logs and writes checkpoints."""
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def log(metrics, step=None):
    if not is_main_process():
        return
    if wandb is not None and getattr(wandb, "run", None) is not None:
        wandb.log(metrics, step=step)


def args_dict(args):
    return {k: (str(v) if isinstance(v, torch.device) else v) for k, v in vars(args).items() if not k.startswith("_")}


def save_checkpoint(checkpoint_dict, is_best, args=None):
    """utils.py:406-419: same dict schema {batch_idx, state_dict, best_loss, optimizer, args}."""
    if not is_main_process():          # ranks hold identical parameters after the all-reduced step: rank 0 writes
        return
    d = run_dir(args)
    checkpoint_file = os.path.join(d, "ckpt.pth.tar")
    best_file = os.path.join(d, "best.pth.tar")
    torch.save(checkpoint_dict, checkpoint_file)
    if is_best:
        shutil.copyfile(checkpoint_file, best_file)


def load_checkpoint(model, optimizer, device, checkpoint_file):
    """utils.py:422-441.  In-place copies keep the optimizer's flat parameter views intact."""
    checkpoint = torch.load(checkpoint_file, map_location=device, weights_only=False)
    with torch.no_grad():
        own = model.state_dict()
        missing = set(own) ^ set(checkpoint["state_dict"])
        if missing:
            raise RuntimeError(f"Error(s) in loading state_dict: key mismatch {sorted(missing)}")
        for k, v in checkpoint["state_dict"].items():
            own[k].copy_(v)
    if optimizer is not None and "optimizer" in checkpoint and not hasattr(optimizer, "_flat"):
        optimizer.load_state_dict(checkpoint["optimizer"])          # utils.py:436 (SGD and other torch optimizers)
    elif optimizer is not None and "optimizer" in checkpoint:
        # FusedAdam: copy the state INTO the flat moment buffers (load_state_dict would replace the views) and
        # restore the hyper-parameters of every group
        sd = checkpoint["optimizer"]
        for grp, saved in zip(optimizer.param_groups, sd.get("param_groups", [])):
            for k, v in saved.items():
                if k != "params":
                    grp[k] = v
        for pid, st in sd.get("state", {}).items():
            p = optimizer.param_groups[0]["params"][pid] if isinstance(pid, int) else None
            if p is None or p not in optimizer.state:
                continue
            for k, v in st.items():
                if isinstance(v, torch.Tensor) and k in optimizer.state[p] and optimizer.state[p][k].shape == v.shape \
                        and optimizer.state[p][k].dim() > 0:
                    optimizer.state[p][k].copy_(v)
                else:
                    optimizer.state[p][k] = v
    print(f"Loaded {checkpoint_file}, trained to epoch {checkpoint['batch_idx']} with best loss (acc for CLIP) "
          f"{checkpoint['best_loss']}")
    return model, optimizer
