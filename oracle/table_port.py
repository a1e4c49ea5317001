"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ld_triangle's table writer.

Only `tests/` may import this module; the product path (`ld_tools_b200`) never does.

Parity status: PINNED through the driver goldens -- `tests/golden/drivers/triangle_*` holds the .tsv files the
unmodified reference `ld_triangle.py` wrote in this container (`tests/golden/make_driver_golden.py`);
`tests/test_oracle.py::test_table_port_reproduces_reference_tables` parses their cells back into the Python
objects they print (int 0 / float) and re-creates the body lines with this port, byte for byte.

It keeps the reference's data structures (a list of lists of Python objects, `str()` per cell): the cell text
is whatever Python prints for the int `0` or the rounded float that calc_ld returned.  Paths are relative to
/root/reference.
"""

__all__ = ["matrix_body"]


def matrix_body(value_of, v, rs_ids, poss, ld_low_thres=None):
    """Body lines of the table (one per variant) as one str.

    value_of(row, col) -> trg_vals[ld_measure] for row > col: the rounded measure calc_ld returned for
    var_1 = row variant, var_2 = column variant (ld_triangle.py:193), an int 0 or a float.
    rs_ids / poss: the sorted identifiers and positions (poss already str, ld_triangle.py:353).
    """
    ld_two_dim = [[0 for col_index in range(v)] for row_index in range(v)]      # :114
    for row_index in range(v):                                                  # :133-134
        for col_index in range(v):
            if row_index <= col_index:                                          # :150
                continue
            val = value_of(row_index, col_index)
            if ld_low_thres is not None:                                        # :223-225
                if val < ld_low_thres:
                    continue
            ld_two_dim[row_index][col_index] = val                              # :230
    out = []
    for row_index in range(v):                                                  # :356-360
        line = '\t'.join(map(str, ld_two_dim[row_index])) + '\n'
        out.append(rs_ids[row_index] + '\t' + poss[row_index] + '\t' + line)
    return ''.join(out)
