"""Alias: reference import path ``llava.eval.eval_spatial`` -> spatialrgpt_b200.eval_spatial (`python -m llava.eval.eval_spatial ...`)."""
from spatialrgpt_b200.eval_spatial import *  # noqa: F401,F403
from spatialrgpt_b200.eval_spatial import build_arg_parser, eval_model  # noqa: F401

if __name__ == "__main__":
    print(f"wrote {eval_model(build_arg_parser().parse_args())} answers")
