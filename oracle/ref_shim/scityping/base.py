from dataclasses import dataclass  # noqa: F401
