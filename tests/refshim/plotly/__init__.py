"""TEST INFRASTRUCTURE ONLY: ld_triangle.py imports plotly unconditionally (ld_triangle.py:378-379);
the golden runs use `-o table`, which never touches it."""
