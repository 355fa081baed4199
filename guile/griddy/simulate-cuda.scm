;;; (griddy simulate-cuda) — drop-in for (griddy simulate): same `simulate' signature
;;; (griddy/simulate.scm:133-134), with the body of `update' (simulate.scm:141-153) replaced by one
;;; call into libgriddy_cuda.so.  A user script switches by importing this module instead:
;;;
;;;   (use-modules (griddy core) (griddy simulate-cuda))
;;;   (simulate make-skeleton add-actors! #:duration 60)
;;;
;;; `simulate/headless' is the fixed-tick-count driver BASELINE.json's configs ask for; the stock loop
;;; cannot run without a window (event-loop.scm:192-211).
;;;
;;; Not executed under Guile here (no Guile in the build image); executed under the fixtures' Scheme
;;; evaluator by tests/test_guile_glue.py: the reference's example scripts, with (griddy simulate) replaced by
;;; this module in their use-modules, load / update / draw through it.
(define-module (griddy simulate-cuda)
  #:use-module (griddy constants)
  #:use-module (griddy math)
  #:use-module (griddy core)
  #:use-module (griddy draw)
  #:use-module (griddy event-loop)
  #:use-module (griddy cuda)
  #:use-module (chickadee graphics path)
  #:use-module (chickadee math vector)
  #:use-module (rnrs bytevectors)
  #:export (simulate simulate/headless))

(define* (simulate make-skeleton add-actors! #:key (width 500) (height 500) (duration #f))
  (define ctx #f)
  (define handle #f)
  (define skeleton-canvas #f)
  (define sim-t0 0)

  (define (load)
    (set! sim-t0 (elapsed-time))
    (let ((world (make-skeleton)))
      (add-actors! world)
      (set! skeleton-canvas (make-skeleton-canvas world))
      (set! ctx (cuda-open))
      (set! handle (cuda-upload-world! ctx world))))

  (define (update delta-t)
    (when (and duration (> (- (elapsed-time) sim-t0) duration))
      (quit))
    (cuda-step! ctx 1))                                   ; was: rebuild world, for-each advance$

  ;; was: (draw-canvas (make-actors-canvas world)) — one (get-pos actor) per actor on the host
  ;; (draw.scm:74-88).  The device computes the positions; the painter is draw.scm:74-76's.
  (define (draw alpha)
    (draw-canvas skeleton-canvas)
    (call-with-values (lambda () (cuda-positions ctx))
      (lambda (n xy)
        (draw-canvas
         (make-canvas
          (with-style ((fill-color *actor/color*))
            (apply superimpose
                   (let loop ((k 0) (acc '()))
                     (if (= k n)
                         acc
                         (let ((x (bytevector-ieee-single-native-ref xy (* 8 k)))
                               (y (bytevector-ieee-single-native-ref xy (+ 4 (* 8 k)))))
                           (loop (1+ k)
                                 (if (nan? x) acc
                                     (cons (fill (circle (vec2 x y) *actor/size*)) acc)))))))))))))

  (define (quit)
    (pk 'last-update-ms (cuda-last-step-ms ctx))
    (cuda-close ctx)
    (abort-game))

  (run-game #:load load #:update update #:update-hz (recip *simulate/time-step*)
            #:draw draw #:quit quit #:window-width width #:window-height height))

(define* (simulate/headless make-skeleton add-actors! #:key (ticks 1000))
  "Advance TICKS ticks without a window and return the resulting world."
  (let ((world (make-skeleton)))
    (add-actors! world)
    (let* ((ctx (cuda-open))
           (handle (cuda-upload-world! ctx world)))
      (cuda-step! ctx ticks)
      (let ((result (cuda-download-world! ctx handle make-skeleton)))
        (cuda-close ctx)
        result))))
