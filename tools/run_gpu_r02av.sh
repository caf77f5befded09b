#!/bin/bash
# round 2, call av: the plain-C consumer tests (incl. the torus loop added after the last full run of the suite)
timeout 200 python -m pytest tests/test_c_consumer.py -q -m gpu 2>&1 | tail -3
