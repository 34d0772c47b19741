"""TEST INFRASTRUCTURE — not part of the product.

CPU restatements of the reference hot path (ConvMixer / MlpMixer forward,
backward, MPJPE loss, Adam).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker / the timed CPU baseline.  The product
package ``motionmixerconv_b200`` never imports it.
"""
