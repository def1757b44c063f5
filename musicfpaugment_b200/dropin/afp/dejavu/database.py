"""afp/dejavu/database.py: the abstract base of the reference is not needed by the in-memory drop-in; the name is
kept importable (`from dejavu.database import BaseDatabase`, postgres_database.py:6)."""
import abc


class BaseDatabase(object, metaclass=abc.ABCMeta):
    type = None

    def __init__(self):
        super().__init__()
