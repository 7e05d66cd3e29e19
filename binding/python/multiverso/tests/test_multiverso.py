"""The reference binding's test-suite (binding/python/multiverso/tests/test_multiverso.py:24-107)
on the compat package: array 10000 x 100 iterations, matrix 11 x 10 with row ops, shared variables
(torch instead of Theano). Works on either backend, 1 process or several (mvrun / torchrun)."""
import os
import sys
import unittest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np

import multiverso as mv


def setUpModule():
    mv.init()


def tearDownModule():
    mv.shutdown()


class TestMultiversoTables(unittest.TestCase):
    def _test_array(self, size):
        tbh = mv.ArrayTableHandler(size)
        mv.barrier()
        for i in range(20):
            tbh.add(range(1, size + 1), sync=True)
            tbh.add(range(1, size + 1), sync=True)
            mv.barrier()
            for j, actual in enumerate(tbh.get()):
                self.assertEqual((j + 1) * (i + 1) * 2 * mv.workers_num(), actual)
            mv.barrier()

    def test_small_array(self):
        self._test_array(1)          # upstream issue #69 is fixed here

    def test_array(self):
        self._test_array(10000)

    def test_matrix(self):
        num_row, num_col = 11, 10
        size = num_col * num_row
        workers_num = mv.workers_num()
        tbh = mv.MatrixTableHandler(num_row, num_col)
        mv.barrier()
        for count in range(1, 11):
            row_ids = [0, 1, 5, 10]
            tbh.add(range(size), sync=True)
            tbh.add([range(rid * num_col, (1 + rid) * num_col) for rid in row_ids], row_ids, sync=True)
            mv.barrier()
            data = tbh.get()
            mv.barrier()
            for i, row in enumerate(data):
                for j, actual in enumerate(row):
                    expected = (i * num_col + j) * count * workers_num
                    if i in row_ids:
                        expected += (i * num_col + j) * count * workers_num
                    self.assertEqual(expected, actual)
            data = tbh.get(row_ids)
            mv.barrier()
            for i, row in enumerate(data):
                for j, actual in enumerate(row):
                    expected = (row_ids[i] * num_col + j) * count * workers_num * 2
                    self.assertEqual(expected, actual)


class TestMultiversoSharedVariable(unittest.TestCase):
    def test_shared_variable(self):
        import torch
        from multiverso.torch_ext import mv_shared, sync_all_mv_shared_vars
        row, col = 200, 200
        W = mv_shared(torch.zeros(row, col))
        delta = torch.arange(row * col, dtype=torch.float32).view(row, col)
        t = W.get_value()
        t += delta
        sync_all_mv_shared_vars()
        mv.barrier()
        sync_all_mv_shared_vars()
        got = W.get_value().cpu()
        self.assertTrue(torch.equal(got, delta * mv.workers_num()))


if __name__ == "__main__":
    unittest.main()
