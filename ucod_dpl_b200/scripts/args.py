"""Command line of the reference launchers (scripts/args.py:5-21) plus the two paths that are hard-coded there."""
import argparse


def parse_train_args(argv=None):
    p = argparse.ArgumentParser(description="UCOD-DPL on B200")
    p.add_argument("--config", required=True, help="config file path")
    p.add_argument("--work_dir", type=str, default="work_dir", help="work dir")
    p.add_argument("--resume", type=str, default=None, help="resume from checkpoint")
    p.add_argument("--load_from", type=str, default=None, help="load from checkpoint")
    p.add_argument("--refiner_path", type=str, default=None, help="load refiner checkpoint")
    p.add_argument("--launcher", choices=["none", "pytorch", "slurm", "mpi"], default="none", help="job launcher")
    p.add_argument("--local_rank", "--local-rank", type=int, default=0)
    # not in the reference CLI (it reads them from the config only)
    p.add_argument("--dataset_dir", type=str, default=None, help="overrides cfg.dataset_cfg.dataset_dir")
    p.add_argument("--datasets", type=str, default=None, help="comma separated test sets (default: the four of eval.py)")
    p.add_argument("--batch_size", type=int, default=64, help="images per launch sequence and GPU")
    p.add_argument("--exp_name", type=str, default=None)
    p.add_argument("--no_save", action="store_true", help="skip the PNG output")
    p.add_argument("--cache_dir", type=str, default=None, help="overrides cfg.dataset_cfg.cache_dir")
    p.add_argument("--max_epoch", type=int, default=None, help="overrides cfg.train_cfg.max_epoch")
    p.add_argument("--no_val", action="store_true", help="train without the periodic Look-Twice validation")
    return p.parse_args(argv)
