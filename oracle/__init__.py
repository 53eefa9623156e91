"""CPU oracle for the deepEMIA post-head hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this
package; nothing under ``deepemia_b200/`` does.  Each function restates the reference's algorithm (numpy / OpenCV /
scipy, exactly the third-party calls the reference makes) and cites the reference file:line it follows.  The restatement
is pinned against the UNMODIFIED reference imported from /root/reference in the build container: see
``tests/golden/make_golden.py`` and the committed vectors under ``tests/golden/*.npz``.
"""
