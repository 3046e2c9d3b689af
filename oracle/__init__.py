"""Test infrastructure (CPU oracle).  See vallex_oracle.py / vocos_oracle.py headers."""
