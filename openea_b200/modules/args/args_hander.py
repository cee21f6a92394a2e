"""Experiment arguments: a flat JSON file becomes an attribute bag.

Interface of the reference's modules/args/args_hander.py (load_args :4, ARGs :13, check_args :19).
"""
import json


class ARGs:
    """Attribute bag over a dict (every JSON key becomes an attribute)."""

    def __init__(self, dic):
        self.__dict__.update(dic)

    def __repr__(self):
        return "ARGs(%r)" % (self.__dict__,)


def load_args(file_path):
    with open(file_path, "r") as fh:
        args_dict = json.load(fh)
    print("load arguments:", args_dict)
    return ARGs(args_dict)


def check_args(args):
    # translation-family models pair every positive with exactly one negative (margin loss broadcasting)
    if getattr(args, "embedding_module", None) in ("TransE", "TransH", "TransR", "TransD"):
        assert args.neg_triple_num == 1
