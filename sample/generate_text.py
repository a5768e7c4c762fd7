"""python -m sample.generate_text ... -- same entry point and flags as the reference's sample/generate_text.py, hot path on B200."""
from surfd_b200.cli import main

if __name__ == "__main__":
    main("text")
