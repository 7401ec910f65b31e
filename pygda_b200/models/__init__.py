"""``pygda.models`` estimators on the accelerated path."""
from .base import BaseGDA
from .a2gnn import A2GNN

__all__ = ["BaseGDA", "A2GNN"]
