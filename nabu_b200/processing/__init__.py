"""On-disk data format and input pipeline of nabu: TFRecord framing and
tf.train.Example parsing without TensorFlow, the audio / text readers, bucketed batching."""
