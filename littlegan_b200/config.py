"""Layered configuration, same schema and precedence as the reference's `config.py:6-42`:
`sample.config.json` <- `<env>.config.json` <- command line.  The three keys at the end of the
packaged sample file (`dtype`, `cuda_graph`, `seed`) and `augment` (the reference's input augmentation,
always on there) are new and default, so reference config files keep loading unchanged.
"""
import json
import os
from argparse import ArgumentParser

_PKG_SAMPLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sample.config.json")
MODES = ["train", "plot", "visual", "random-sample", "evaluate", "condition-sample", "evaluate-sample",
         "export-model"]


class Arg:
    """Attribute bag.  `Arg()` parses sys.argv like the reference; `Arg.from_dict(...)` builds
    one programmatically (tests, bench) on top of the packaged sample config."""

    def __init__(self, argv=None, config_dir="."):
        print(" - Initializing Application...")
        parser = ArgumentParser(prog="LittleGAN", description="The code for paper: LittleGAN")
        parser.add_argument("mode", type=str, help="run mode", default="train", choices=MODES)
        parser.add_argument("exp_name", type=str, help="experience name")
        parser.add_argument("-e", "--env", type=str, help="config environment", default="sample")
        parser.add_argument("-g", "--gpu", type=str, required=False, help="gpu ids, eg: 0,1,2,3", default="-1")
        parser.add_argument("--debug", help="use debug mode, ignore git repo is dirty", action="store_true")
        cli = parser.parse_args(argv)
        sample = os.path.join(config_dir, "sample.config.json")
        self._load(sample if os.path.isfile(sample) else _PKG_SAMPLE)
        self.env_file = cli.env + ".config.json"
        env_path = os.path.join(config_dir, self.env_file)
        if cli.env != "sample" or os.path.isfile(env_path):
            self._load(env_path)
        for item, value in vars(cli).items():
            setattr(self, item, value)
        self._derive()

    def _load(self, path):
        with open(path) as f:
            for item, value in json.load(f).items():
                setattr(self, item, value)

    def _derive(self):
        # config.py:32-39
        if self.attr is None:          # the reference crashes here; 40 = all CelebA attributes
            self.attr = list(range(40))
        self.cond_dim = len(self.attr)
        self.result_dir = os.path.join(self.all_result_dir, getattr(self, "exp_name", "exp"))
        gpu = getattr(self, "gpu", "-1")
        if isinstance(gpu, str):
            if gpu != "-1":
                os.environ["CUDA_VISIBLE_DEVICES"] = gpu
            gpu = [int(i) for i in gpu.split(",") if i.isnumeric() and int(i) >= 0]
        self.gpu = gpu
        self.prefetch = self.prefetch_batch * self.batch_size
        for key, default in (("dtype", "bf16"), ("cuda_graph", True), ("seed", 0), ("augment", True)):
            if not hasattr(self, key):
                setattr(self, key, default)

    @classmethod
    def from_dict(cls, **overrides):
        self = cls.__new__(cls)
        self._load(_PKG_SAMPLE)
        self.mode, self.exp_name, self.env, self.debug, self.gpu = "train", "exp", "sample", True, []
        self.env_file = "sample.config.json"
        for k, v in overrides.items():
            setattr(self, k, v)
        self._derive()
        return self

    def __str__(self):
        return self.__dict__.__str__()
