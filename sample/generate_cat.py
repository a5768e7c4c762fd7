"""python -m sample.generate_cat ... -- same entry point and flags as the reference's sample/generate_cat.py, hot path on B200."""
from surfd_b200.cli import main

if __name__ == "__main__":
    main("cat")
