"""configargparse stand-in: argparse plus ``is_config_file`` arguments.

The reference's parsers (NP/run_nerf.py:424, NP/run_nerf_view.py:672) declare ``--config`` with ``is_config_file=True`` and
read files of ``key = value`` lines (NP/configs*/**.txt); command-line values override file values, file values override
defaults; a flag (store_true) is set by ``key = True``.
"""
import argparse
import sys

__version__ = "0-cnerf-shim"


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *args, **kwargs):
        kwargs.pop("default_config_files", None)
        kwargs.pop("config_file_parser_class", None)
        super().__init__(*args, **kwargs)
        self._config_dests = []

    def add_argument(self, *names, **kwargs):
        is_cfg = kwargs.pop("is_config_file", False)
        kwargs.pop("env_var", None)
        action = super().add_argument(*names, **kwargs)
        if is_cfg:
            self._config_dests.append(action.dest)
        return action

    add = add_argument

    @staticmethod
    def _read_config(path):
        items = []
        with open(path) as f:
            for line in f:
                line = line.split("#", 1)[0].strip()
                if not line or line.startswith(";") or line.startswith("["):
                    continue
                key, sep, val = line.partition("=")
                if not sep:
                    key, _, val = line.partition(":")
                items.append((key.strip(), val.strip().strip('"').strip("'")))
        return items

    def parse_known_args(self, args=None, namespace=None):
        args = list(sys.argv[1:] if args is None else args)
        # find config files named on the command line
        opt_of_dest = {a.dest: a for a in self._actions}
        cfg_flags = [s for d in self._config_dests for s in opt_of_dest[d].option_strings]
        cfg_paths = []
        for i, tok in enumerate(args):
            for flag in cfg_flags:
                if tok == flag and i + 1 < len(args):
                    cfg_paths.append(args[i + 1])
                elif tok.startswith(flag + "="):
                    cfg_paths.append(tok[len(flag) + 1:])
        file_args = []
        by_name = {}
        for a in self._actions:
            for s in a.option_strings:
                by_name[s.lstrip("-")] = a
        for path in cfg_paths:
            for key, val in self._read_config(path):
                a = by_name.get(key)
                if a is None:
                    continue                                     # unknown keys are ignored, like configargparse's default
                flag = a.option_strings[0]
                if isinstance(a, (argparse._StoreTrueAction, argparse._StoreFalseAction, argparse._StoreConstAction)):
                    if val.lower() in ("true", "yes", "1", "on"):
                        file_args.append(flag)
                elif a.nargs in ("+", "*") or isinstance(a.nargs, int):
                    file_args += [flag] + val.strip("[]").replace(",", " ").split()
                else:
                    file_args += [flag, val]
        return super().parse_known_args(file_args + args, namespace)


ArgParser = ArgumentParser
Namespace = argparse.Namespace
