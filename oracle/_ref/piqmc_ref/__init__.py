# built by oracle/build_ref.py from the reference's own .pyx files
