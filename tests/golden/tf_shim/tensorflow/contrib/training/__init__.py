"""tf.contrib.training namespace; HParams is attached by tensorflow/__init__.py."""
HParams = None
