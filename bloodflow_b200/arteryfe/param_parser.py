"""ParamParser -- same files, same quirks as the reference's
arteryfe/param_parser.py (argparse ``--cfg``; three sections; ``eval`` of the
values; a comma in [Parameters] makes a float array), rewritten for Python 3.12
(``SafeConfigParser`` is gone).  ``ParamParser(path)`` is also accepted, as the
reference's tutorial shows (docs/source/tutorial.rst:15)."""
import argparse
import configparser
import sys

import numpy as np


class ParamParser(object):

    def __init__(self, fparams=None):
        if fparams is None:
            ap = argparse.ArgumentParser()
            ap.add_argument("--cfg")
            fparams = ap.parse_known_args()[0].cfg
        self.fparams = fparams
        try:
            with open(self.fparams, 'r') as fh:
                fh.read()
        except Exception as e:              # param_parser.py:16-22
            print(e)
            sys.exit(1)
        self.param, self.geo, self.solution = self.get_params()

    def get_params(self):
        cfg = configparser.ConfigParser()
        cfg.optionxform = str               # keep the case of the keys
        cfg.read(self.fparams)
        return (self.get_param_section(cfg), self.get_section(cfg, "Geometry"),
                self.get_section(cfg, "Solution"))

    @staticmethod
    def get_param_section(config):
        out = {}
        for key, text in config.items('Parameters'):
            if "," in text:
                out[key] = np.array([float(tok) for tok in text.split(',')])
            else:
                out[key] = eval(text)
        return out

    @staticmethod
    def get_section(config, section):
        out = {}
        for key, text in config.items(section):
            try:
                out[key] = eval(text)
            except Exception:
                out[key] = text
        return out
