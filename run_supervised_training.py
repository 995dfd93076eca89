#!/usr/bin/env python
"""Supervised wavefunction optimisation driver: the reference's
run_supervised_training.py (lines 73-148) re-hosted on the B200 path."""
import os
import sys

from absl import app
from absl import flags

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cgs_vmc_b200 import checkpoint, drivers, training, utils, wavefunctions  # noqa: E402
from cgs_vmc_b200.session import Session  # noqa: E402

flags.DEFINE_string('checkpoint_dir', '', 'Full path to the checkpoint directory.')
flags.DEFINE_string('supervisor_dir', '', 'Directory with the supervisor checkpoints.')
flags.DEFINE_integer('num_epochs', 1000, 'Total of number of epochs to train on.')
flags.DEFINE_integer('checkpoint_frequency', 1, 'Number of epochs between checkpoints.')
flags.DEFINE_boolean('resume_training', False, 'Restore variables from the latest checkpoint.')
flags.DEFINE_string('wavefunction_type', 'fully_connected', 'Key of wavefunctions.WAVEFUNCTION_TYPES.')
flags.DEFINE_string('optimizer', 'SWO', 'Key of training.SUPERVISED_OPTIMIZERS.')
flags.DEFINE_string('list_of_evaluators', '', 'Unused.')
flags.DEFINE_boolean('generate_vectors', False, 'Not available on this path.')
flags.DEFINE_string('basis_file_path', '', 'Unused.')
flags.DEFINE_string('hparams', '', 'Comma-separated name=value overrides.')
flags.DEFINE_boolean('override', True, 'Whether to override an existing hparams file.')
FLAGS = flags.FLAGS


def main(argv):
  del argv
  rank = drivers.init_distributed()
  supervisor_hparams = utils.load_hparams(os.path.join(FLAGS.supervisor_dir, 'hparams.pbtxt'))
  hparams = utils.create_hparams()
  hparams.set_hparam('num_sites', supervisor_hparams.num_sites)
  hparams.set_hparam('size_x', supervisor_hparams.size_x)
  hparams.set_hparam('size_y', supervisor_hparams.size_y)
  hparams.set_hparam('checkpoint_dir', FLAGS.checkpoint_dir)
  hparams.set_hparam('supervisor_dir', FLAGS.supervisor_dir)
  hparams.set_hparam('basis_file_path', FLAGS.basis_file_path)
  hparams.set_hparam('num_epochs', FLAGS.num_epochs)
  hparams.set_hparam('wavefunction_type', FLAGS.wavefunction_type)
  hparams.parse(FLAGS.hparams)
  hparams_path = os.path.join(hparams.checkpoint_dir, 'hparams.pbtxt')
  if rank == 0:
    os.makedirs(FLAGS.checkpoint_dir, exist_ok=True)
    if os.path.exists(hparams_path) and not FLAGS.override:
      print('Hparams file already exists')
      sys.exit()
    utils.save_hparams(hparams, hparams_path)

  target_wavefunction = wavefunctions.build_wavefunction(supervisor_hparams)
  wavefunction = wavefunctions.build_wavefunction(hparams)
  wavefunction_optimizer = training.SUPERVISED_OPTIMIZERS[FLAGS.optimizer]()
  shared_resources = {}
  train_ops = wavefunction_optimizer.build_opt_ops(
      wavefunction=wavefunction, target_wavefunction=target_wavefunction, hparams=hparams,
      shared_resources=shared_resources)

  session = Session()
  target_saver = checkpoint.Saver(target_wavefunction)
  target_saver.restore(session, checkpoint.latest_checkpoint(FLAGS.supervisor_dir))
  checkpoint_saver = checkpoint.Saver(wavefunction, max_to_keep=5)
  if FLAGS.resume_training:
    checkpoint_saver.restore(session, checkpoint.latest_checkpoint(hparams.checkpoint_dir))

  metrics_path = os.path.join(hparams.checkpoint_dir, 'metrics.txt')
  for epoch_number in range(FLAGS.num_epochs):
    wavefunction_optimizer.run_optimization_epoch(train_ops, session, hparams, epoch_number)
    if rank == 0:
      with open(metrics_path, 'a') as f:
        f.write('{}\n'.format(session.run(train_ops.metrics)))
      if epoch_number % FLAGS.checkpoint_frequency == 0:
        name = 'model_after_{}_epochs'.format(epoch_number)
        checkpoint_saver.save(session, os.path.join(hparams.checkpoint_dir, name))
    else:
      session.run(train_ops.metrics)        # keeps the collective in step


if __name__ == '__main__':
  app.run(main)
