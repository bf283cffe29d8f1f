"""Config-driven entry -- the `graphvite run <config.yaml>` part of the reference's CLI
(python/graphvite/cmd.py:82-163) for the node-embedding application:

    python -m graphvite_b200.cmd run config/graph/line_youtube.yaml [--no-eval] [--gpu N] [--cpu N] [--epoch N]

The YAML sections and keys are the reference's (doc/source/user/configuration.rst): application,
resource {gpus, cpu_per_gpu, dim, gpu_memory_limit}, format, graph, build, load, train, evaluate,
save.  "auto" maps to the auto sentinel; "<dataset.split>" placeholders need the reference's dataset
downloader and are rejected here (no network) -- give file names instead.
"""
import argparse
import logging
import re
import sys

import yaml

from . import application as gap
from . import optimizer as _optimizer
from .base import auto

logger = logging.getLogger(__name__)


def _map(value, function):
    if isinstance(value, dict):
        return {k: _map(v, function) for k, v in value.items()}
    if isinstance(value, list):
        return [_map(v, function) for v in value]
    return function(value)


def load_config(config_file):
    """cmd.py:82-106"""
    with open(config_file, "r") as fin:
        cfg = yaml.safe_load(fin)

    def resolve(x):
        if x == "auto":
            return auto
        if isinstance(x, str) and re.match(r"<(\w+)\.(\w+)>", x):
            raise ValueError("dataset placeholder `%s` needs the reference's downloader; put a file name in the "
                             "configuration instead" % x)
        return x

    cfg = _map(cfg, resolve)
    if "optimizer" in cfg.get("build", {}) and isinstance(cfg["build"]["optimizer"], dict):
        cfg["build"]["optimizer"] = _optimizer.Optimizer(**cfg["build"]["optimizer"])
    return cfg


def run_main(args):
    """cmd.py:140-163"""
    cfg = load_config(args.config)
    resource = dict(cfg.get("resource", {}))
    if args.gpu is not None:
        resource["gpus"] = list(range(args.gpu))
    if args.cpu is not None:
        resource["cpu_per_gpu"] = args.cpu
    train = dict(cfg.get("train", {}))
    if args.epoch is not None:
        train["num_epoch"] = args.epoch
    app = gap.Application(cfg["application"], **resource)
    if "format" in cfg:
        app.set_format(**cfg["format"])
    app.load(**cfg["graph"])
    app.build(**cfg.get("build", {}))
    if "load" in cfg:
        app.load_model(**cfg["load"])
    app.train(**train)
    results = []
    if args.eval and "evaluate" in cfg:
        evaluations = cfg["evaluate"] if isinstance(cfg["evaluate"], list) else [cfg["evaluate"]]
        for evaluation in evaluations:
            results.append(app.evaluate(**evaluation))
    if "save" in cfg:
        app.save_model(**cfg["save"])
    return app, results


def get_parser():
    parser = argparse.ArgumentParser(description="graphvite_b200 command line interface")
    command = parser.add_subparsers(metavar="command", dest="command")
    command.required = True
    run = command.add_parser("run", help="run from configuration file")
    run.add_argument("config", help="yaml configuration file")
    run.add_argument("--no-eval", help="turn off evaluation", dest="eval", action="store_false")
    run.add_argument("--gpu", help="override the number of GPUs", type=int)
    run.add_argument("--cpu", help="override the number of CPUs per GPU", type=int)
    run.add_argument("--epoch", help="override the number of epochs", type=int)
    return parser


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    args = get_parser().parse_args(argv)
    if args.command == "run":
        run_main(args)


if __name__ == "__main__":
    main(sys.argv[1:])
