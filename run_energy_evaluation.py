#!/usr/bin/env python
"""Energy evaluation driver: the reference's run_energy_evaluation.py
(lines 42-91) re-hosted on the B200 path."""
import os
import sys

import numpy as np
from absl import app
from absl import flags

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cgs_vmc_b200 import checkpoint, drivers, evaluation, operators, utils, wavefunctions  # noqa: E402
from cgs_vmc_b200.session import Session  # noqa: E402

flags.DEFINE_float('heisenberg_jx', 1.0, 'Jx value in Heisenberg Hamiltonian.')
flags.DEFINE_string('checkpoint_dir', '', 'Full path to the checkpoint directory.')
flags.DEFINE_string('output_file', '', 'Optional file for the result.')
flags.DEFINE_string('hparams', '', 'Comma-separated name=value overrides.')
FLAGS = flags.FLAGS


def main(argv):
  del argv
  rank = drivers.init_distributed()
  hparams = utils.load_hparams(os.path.join(FLAGS.checkpoint_dir, 'hparams.pbtxt'))
  hparams.parse(FLAGS.hparams)
  bonds, j_x, j_z = drivers.load_bonds(FLAGS.checkpoint_dir, hparams.num_sites, FLAGS.heisenberg_jx)
  wavefunction = wavefunctions.build_wavefunction(hparams)
  hamiltonian = operators.HeisenbergHamiltonian(bonds, j_x, j_z)
  evaluator = evaluation.MonteCarloOperatorEvaluator()
  shared_resources = {}
  evaluation_ops = evaluator.build_eval_ops(
      wavefunction=wavefunction, operator=hamiltonian, hparams=hparams,
      shared_resources=shared_resources)
  session = Session()
  checkpoint.Saver(wavefunction).restore(
      session, checkpoint.latest_checkpoint(hparams.checkpoint_dir))
  data = evaluator.run_evaluation(evaluation_ops, session, hparams, epoch_num=0)
  mean_energy = np.mean(data)
  # the reference prints sqrt(std) / n (run_energy_evaluation.py:87, SURVEY.md
  # appendix B-11); the standard error of the mean is reported next to it
  reference_uncertainty = np.sqrt(np.std(data)) / len(data)
  stderr = np.std(data) / np.sqrt(len(data))
  if rank == 0:
    print('Energy: {} +/- {}'.format(mean_energy, reference_uncertainty))
    print('Standard error of the mean: {}'.format(stderr))
    if FLAGS.output_file:
      with open(FLAGS.output_file, 'w') as f:
        f.write('Energy mean: {}\nUncertainty: {}\nStderr: {}\n'.format(
            mean_energy, reference_uncertainty, stderr))


if __name__ == '__main__':
  app.run(main)
