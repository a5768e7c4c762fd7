"""python -m sample.generate_sketch ... -- same entry point and flags as the reference's sample/generate_sketch.py, hot path on B200."""
from surfd_b200.cli import main

if __name__ == "__main__":
    main("sketch")
