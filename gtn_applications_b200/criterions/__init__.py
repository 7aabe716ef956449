"""Drop-in counterparts of the reference's criterions package
(criterions/{ctc,asg,stc,transducer}.py), computed by libwfst_b200.so."""
