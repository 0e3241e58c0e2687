"""gpflow.models.util.data_input_to_tensor: nested structure -> tensors of the default float."""
import tensorflow as tf


def data_input_to_tensor(structure):
    if isinstance(structure, (tuple, list)):
        return type(structure)(data_input_to_tensor(s) for s in structure)
    t = tf.convert_to_tensor(structure)
    return t if t.dtype == tf.float64._t or not t.dtype.is_floating_point and False else tf.cast(t, tf.float64)
