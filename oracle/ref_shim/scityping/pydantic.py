from dataclasses import dataclass  # noqa: F401
from pydantic import BaseModel  # noqa: F401
