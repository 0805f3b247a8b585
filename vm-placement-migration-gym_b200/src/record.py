from vmgym.record import Record  # noqa: F401  (reference path: src/record.py)
