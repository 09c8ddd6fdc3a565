"""Alias: reference import path ``llava.eval.eval_region_cls`` -> spatialrgpt_b200.eval_region_cls (`python -m llava.eval.eval_region_cls ...`)."""
from spatialrgpt_b200.eval_region_cls import *  # noqa: F401,F403
from spatialrgpt_b200.eval_region_cls import build_arg_parser, eval_model  # noqa: F401

if __name__ == "__main__":
    _a = build_arg_parser().parse_args()
    print(f"wrote {eval_model(_a, seed=_a.seed)} answers")
