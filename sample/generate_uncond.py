"""python -m sample.generate_uncond ... -- same entry point and flags as the reference's sample/generate_uncond.py, hot path on B200."""
from surfd_b200.cli import main

if __name__ == "__main__":
    main("uncond")
