"""utils/parser_util.py of the reference: generate_args() (:164-170) with the same flags"""
from ...cli import generate_args  # noqa: F401
