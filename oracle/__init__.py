"""CPU oracle for the UP-Retinex classical hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package.  The product path (``retinex_image_enhancement_b200``) never does.
"""
