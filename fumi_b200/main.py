"""CLI entry: mirror of the reference's fumi/main.py (same flags, seed order and dispatch).

    python -m fumi_b200.main --model fumi --num_shots 5 --synthetic --wandb_offline
"""
import os
import random
import sys

import numpy as np
import torch

from . import am3, fumi, maml, utils
from .data.loader import get_dataset


def main(args):
    results_path = f"{args.log_dir}/results"
    os.makedirs(results_path, exist_ok=True)
    if utils.wandb is not None and not args.wandb_offline:          # main.py:25-32 (optional here)
        os.environ["WANDB_MODE"] = "online"
        utils.wandb.init(entity=args.wandb_entity, project=args.wandb_project, group=args.wandb_experiment,
                         job_type="eval" if args.evaluate else "train", save_code=True)
        utils.wandb.config.update(args)
    if args.image_embedding_model not in ["resnet-152", "resnet-34"]:
        raise ValueError("Image embedding model must be one of resnet-152 resnet-34")
    if args.image_embedding_model == "resnet-152" and args.im_emb_dim != 2048:
        raise ValueError("Resnet-152 outputs 2048-dimensional embeddings, hence --im_emb_dim should be set to 2048")
    if args.image_embedding_model == "resnet-34" and args.im_emb_dim != 512:
        raise ValueError("Resnet-34 outputs 512-dimensional embeddings, hence --im_emb_dim should be set to 512")

    rank, world = init_distributed(args)
    train_loader, val_loader, test_loader, dictionary = get_dataset(args)
    bs = args.tasks_per_batch or args.batch_size
    args.batch_size = bs
    if world > 1:
        # tasks shard across ranks: every rank draws the SAME global stream of world * bs tasks per
        # step and keeps its slice, so the run samples exactly the tasks of a single-process run with that
        # meta-batch; gradients meet in one NCCL all-reduce (engine._allreduce_*).  Validation / test are replicated.
        train_loader.set_shard(rank, world)
    max_test_batches = int(args.num_ep_test / bs)
    torch.manual_seed(args.seed)                 # main.py:51-53: after the loaders, before the model
    np.random.seed(args.seed)
    random.seed(args.seed)
    model = utils.init_model(args, dictionary)
    if rank == 0:
        print(model)
    optimizer = utils.init_optim(args, model)
    if args.checkpoint:                          # main.py:61-76 restores from wandb; here: a local file
        opt = optimizer[0] if isinstance(optimizer, tuple) else optimizer
        model, _ = utils.load_checkpoint(model, opt, args.device, args.checkpoint)
    if not args.evaluate:
        if args.model == "maml":
            model = maml.training_run(args, model, optimizer, train_loader, val_loader, max_test_batches // 2)
        elif args.model == "fumi":
            model = fumi.training_run(args, model, optimizer, train_loader, val_loader, max_test_batches // 2)
        else:
            model = am3.training_run(args, model, optimizer, train_loader, val_loader, max_test_batches // 2)
    if args.model == "maml":
        test_loss, test_acc = maml.test_loop(args, model, test_loader, max_test_batches)
    elif args.model == "fumi":
        test_loss, test_acc, _, _ = fumi.test_loop(args, model, test_loader, max_test_batches)
    else:
        out = am3.test_loop(args, model, test_loader, max_test_batches)
        test_loss, test_acc = out[0], out[1]
        print(f"test f1: {out[2]}, test prec: {out[3]}, test rec: {out[4]}, test avg lamda: {out[5]}")
    if rank == 0:
        print(f"\n TEST: \ntest loss: {test_loss}, test acc: {test_acc}")
        utils.log({"test/acc": test_acc, "test/loss": test_loss})
    return test_loss, test_acc


def init_distributed(args):
    """One process per GPU under torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment): NCCL
    process group, the rank's device, and rank-0-only logging / checkpoints (utils.is_main_process).  The reference is
    single-process (main.py:145-146); without WORLD_SIZE > 1 nothing here runs."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch.distributed as dist
    if not dist.is_initialized():
        torch.cuda.set_device(args.device)
        dist.init_process_group("nccl", device_id=args.device)
    return dist.get_rank(), dist.get_world_size()


def parse_args(argv=None):
    args = utils.parser().parse_args(sys.argv[1:] if argv is None else argv)
    if args.disable_cuda or not torch.cuda.is_available():
        raise SystemExit("fumi_b200 runs the episodic path on CUDA (sm_100a) only; there is no CPU path")
    args.device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    print(f"running on device {args.device}")
    return args


if __name__ == "__main__":
    main(parse_args())
