"""Table options and per-request options.

Reference: AddOption / GetOption (include/multiverso/updater/updater.h:10-110) and the
``*TableOption`` structs bound to worker/server types by DEFINE_TABLE_TYPE
(include/multiverso/table_interface.h:77-80, table/array_table.h:66-73,
table/matrix_table.h:116-127, table/matrix.h:14-22, table/kv_table.h:120-124).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Optional


class AddOption:
    """{worker_id, momentum, learning_rate, rho, lambda}; defaults {MV_WorkerId(), 0, .01, .1, .1}.

    Field order follows the reference *accessors* (SURVEY Q3), 20-byte wire layout kept."""

    __slots__ = ("worker_id", "momentum", "learning_rate", "rho", "lambda_")

    def __init__(self, worker_id: Optional[int] = None, momentum: float = 0.0,
                 learning_rate: float = 0.01, rho: float = 0.1, lambda_: float = 0.1):
        if worker_id is None:
            from ..runtime import Runtime
            rt = Runtime.get()
            worker_id = rt.worker_id() if rt.started else 0
        self.worker_id = int(worker_id)
        self.momentum = float(momentum)
        self.learning_rate = float(learning_rate)
        self.rho = float(rho)
        self.lambda_ = float(lambda_)

    def pack(self) -> bytes:
        return struct.pack("<iffff", self.worker_id, self.momentum, self.learning_rate, self.rho,
                           self.lambda_)

    @classmethod
    def unpack(cls, data: bytes) -> "AddOption":
        w, m, lr, rho, lam = struct.unpack("<iffff", data[:20])
        return cls(w, m, lr, rho, lam)

    def __repr__(self):
        return (f"AddOption {self.worker_id} {self.momentum} {self.learning_rate} {self.rho} "
                f"{self.lambda_}")


class GetOption:
    __slots__ = ("worker_id",)

    def __init__(self, worker_id: Optional[int] = None):
        if worker_id is None:
            from ..runtime import Runtime
            rt = Runtime.get()
            worker_id = rt.worker_id() if rt.started else 0
        self.worker_id = int(worker_id)

    def pack(self) -> bytes:
        return struct.pack("<i", self.worker_id)


@dataclass
class ArrayTableOption:
    size: int
    dtype: str = "float32"


@dataclass
class MatrixTableOption:
    num_row: int
    num_col: int
    dtype: str = "float32"
    min_value: Optional[float] = None   # server-side random-uniform init (matrix_table.h:102)
    max_value: Optional[float] = None


@dataclass
class MatrixOption(MatrixTableOption):
    is_sparse: bool = False
    is_pipeline: bool = False


@dataclass
class SparseMatrixTableOption(MatrixTableOption):
    is_pipeline: bool = False


@dataclass
class KVTableOption:
    key_dtype: str = "int64"
    val_dtype: str = "float32"
