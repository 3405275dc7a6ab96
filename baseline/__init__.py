"""Reference arm of bench.py: the unmodified reference under baseline/_ref (see install_ref.py)."""
