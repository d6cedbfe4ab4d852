"""Evaluators (reference: nabu/neuralnetworks/evaluators/) -- SURVEY.md section 8 row f2."""
