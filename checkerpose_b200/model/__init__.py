"""Drop-in replacements for checkerpose/model/{backbone,init,init_lm,pipeline,pipeline_lm}.py."""
